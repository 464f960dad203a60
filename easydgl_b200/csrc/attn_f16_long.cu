// attn_f16_long.cu - the fused self-modulating attention core (temporal.py:281-315, 345-385, 412-447) for LONG
// sequences and / or head dim 32 (BASELINE config C5: L = 512, d = 256, h = 8), scaled 3xFP16 on mma.sync.m16n8k16.
//
// attn_f16.cu keeps the whole probability row of a query in registers (L <= 208).  Here the keys are streamed in
// chunks of 64 and the row is never materialised.  The intensity gate G[q,k] = sum_e lam[q,e] marks[k,e] needs the
// FINISHED H[q] = sum_k P[q,k] T[k] (lam is a function of H), so the kernel makes two passes over the keys
// ("two-phase key tiling"):
//
//   pass 1   per chunk: S = Q K_c^T, masks, ONLINE softmax (running row max m and sum l), H <- H alpha + P_c T_c
//   between  H / l -> intensity MLP (sigmoid-dense, per-event dot, softplus) -> lam (A-operand fragments)
//   pass 2   per chunk: S recomputed, P = 2^(S - m) / l, G_c = lam M_c^T (set_diag), O <- O + (G_c o P_c) V_c
//   end      O + residual -> global
//
// One CTA = 16 warps = 256 query rows of one (sequence, head); a warp owns 16 rows in the MMA accumulator layout,
// exactly as in attn_f16.cu, and the MLP section is the same code.  Chunks are staged cooperatively: every thread
// prefetches its share of the NEXT chunk's K and T (or V) rows into registers while the current chunk is computed,
// then the chunk maximum is reduced (the power-of-two scale of the fp16 split is per CHUNK here, applied to the fp32
// accumulators when a chunk's product is folded into the running H / O) and the (hi | lo) fp16 rows are written to
// shared memory in the ldmatrix layout.  Two barriers per chunk.
#include "attn_f16_common.cuh"

namespace edgl {
namespace {

using namespace f16c;

constexpr int KC = 64;        // keys per chunk
constexpr int NTC = KC / 8;   // 8-key tiles per chunk
constexpr int LTHR = 512;     // 16 warps
constexpr int QROWS = 256;    // query rows per CTA

template <int DH>
struct LongLayout {
  using LY = F16Layout<DH>;
  static constexpr int RB = LY::RB;
  static constexpr int O_K = LY::PACK_BYTES;      // [KC] rows of RB bytes: K chunk
  static constexpr int O_X = O_K + KC * RB;       // T chunk (pass 1) / V chunk (pass 2)
  static constexpr int O_M = O_X + KC * RB;       // marks chunk as fp16 fragments, [KC][8] words
  static constexpr int O_RED = O_M + KC * 32;     // unsigned [2][4]: chunk maxima (K, X), double-buffered by parity
  static constexpr int O_KM = O_RED + 32;         // float [Lpad]: min-mask of every key of the sequence
  static size_t bytes(int lpad) { return (size_t)O_KM + (size_t)lpad * 4; }
};

template <int DH>
__global__ void __launch_bounds__(LTHR, 1) attention_f16_long_kernel(AttnArgs a, int qblocks, int nchunks) {
  using LY = F16Layout<DH>;
  using LL = LongLayout<DH>;
  constexpr int E = LY::E, KD = LY::KD, ND = LY::ND, NC = LY::NC, SKW = LY::SKW, RB = LY::RB;
  constexpr int V4 = DH / 4;
  constexpr int KI = (KC * V4 + LTHR - 1) / LTHR;  // (key, 4 dims) items per thread and operand
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const uint32_t* W1t = reinterpret_cast<const uint32_t*>(smem_raw);          // [NC][SKW]
  const float4* bw = reinterpret_cast<const float4*>(smem_raw + LY::BW_OFF);  // [NC/2]
  const float* wv = reinterpret_cast<const float*>(smem_raw + LY::WV_OFF);    // [NC]
  const float* sc = reinterpret_cast<const float*>(smem_raw + LY::SC_OFF);    // [E]
  const float* misc = reinterpret_cast<const float*>(smem_raw + LY::MISC_OFF);
  unsigned char* Ks = smem_raw + LL::O_K;
  unsigned char* Xs = smem_raw + LL::O_X;
  uint32_t* Ms = reinterpret_cast<uint32_t*>(smem_raw + LL::O_M);
  unsigned int* red = reinterpret_cast<unsigned int*>(smem_raw + LL::O_RED);
  float* km = reinterpret_cast<float*>(smem_raw + LL::O_KM);

  const int L = a.L, B = a.B;
  const int qblk = blockIdx.x % qblocks, hh = (blockIdx.x / qblocks) % a.h, b = blockIdx.x / (qblocks * a.h);
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const long long row0 = (long long)b * L;
  const int q0 = qblk * QROWS + warp * 16;
  const bool active = q0 < L;              // warps past the end only help with the staging
  const int qa = q0 + g, qb = q0 + g + 8;  // this thread's two query rows
  const long long ra = row0 + (qa < L ? qa : L - 1), rb = row0 + (qb < L ? qb : L - 1);

  // ---------------------------------------------------------------- constants, key mask, Q fragments
  {
    const unsigned char* src = reinterpret_cast<const unsigned char*>(a.mlp_pack);
    for (int i = tid; i < LY::PACK_BYTES / 16; i += LTHR) cp_async16(smem_raw + i * 16, src + i * 16);
  }
  const int lpad = nchunks * KC;
  // min-mask: +inf real key, fill = masked id (temporal.py:358,425), -inf = beyond L
  for (int k = tid; k < lpad; k += LTHR) km[k] = (k < L) ? (a.kmask[row0 + k] ? INFINITY : kFillMma) : -INFINITY;
  if (tid < 8) red[tid] = 0u;
  __syncthreads();

  float iqa = 1.f, iqb = 1.f, spa = 0.f, spb = 0.f;
  uint32_t qh[KD][4], ql[KD][4];
  if (active) {
    float2 xa[KD][2], xb[KD][2];  // raw Q values of rows g / g+8 (k slots 2t,2t+1 / 2t+8,2t+9 per k16 step)
#pragma unroll
    for (int ks = 0; ks < KD; ++ks)
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        xa[ks][hf] = __ldg(reinterpret_cast<const float2*>(a.Q + ra * a.ldq + hh * DH + ks * 16 + hf * 8 + 2 * t));
        xb[ks][hf] = __ldg(reinterpret_cast<const float2*>(a.Q + rb * a.ldq + hh * DH + ks * 16 + hf * 8 + 2 * t));
      }
    spa = __ldg(a.spans + ra);
    spb = __ldg(a.spans + rb);
    float ma = 0.f, mb = 0.f;
#pragma unroll
    for (int ks = 0; ks < KD; ++ks)
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        ma = fmaxf(ma, fmaxf(fabsf(xa[ks][hf].x), fabsf(xa[ks][hf].y)));
        mb = fmaxf(mb, fmaxf(fabsf(xb[ks][hf].x), fabsf(xb[ks][hf].y)));
      }
    ma = fmaxf(ma, __shfl_xor_sync(0xffffffffu, ma, 1));
    ma = fmaxf(ma, __shfl_xor_sync(0xffffffffu, ma, 2));
    mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, 1));
    mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, 2));
    float sqa, sqb;
    pow2_scale(ma, sqa, iqa);
    pow2_scale(mb, sqb, iqb);
#pragma unroll
    for (int ks = 0; ks < KD; ++ks) {
      split2(xa[ks][0].x * sqa, xa[ks][0].y * sqa, qh[ks][0], ql[ks][0]);
      split2(xb[ks][0].x * sqb, xb[ks][0].y * sqb, qh[ks][1], ql[ks][1]);
      split2(xa[ks][1].x * sqa, xa[ks][1].y * sqa, qh[ks][2], ql[ks][2]);
      split2(xb[ks][1].x * sqb, xb[ks][1].y * sqb, qh[ks][3], ql[ks][3]);
    }
  }

  // this lane's ldmatrix row addresses (matrix m = lane / 8, row r = lane % 8), as in attn_f16.cu
  const int lm = lane >> 3, lr = lane & 7;
  const uint32_t ks_addr = (uint32_t)__cvta_generic_to_shared(Ks) + lr * RB + (lm & 1) * 16 + (lm >> 1) * (DH * 2);
  const uint32_t xs_addr = (uint32_t)__cvta_generic_to_shared(Xs) + ((lm & 1) * 8 + lr) * RB + (lm >> 1) * 16;

  // ---------------------------------------------------------------- chunk staging
  float4 kreg[KI], xreg[KI];
  uint4 mreg = make_uint4(0u, 0u, 0u, 0u);
  auto load_chunk = [&](int c, const float* X, int ldx, bool want_marks) {
#pragma unroll
    for (int it = 0; it < KI; ++it) {
      const int i = tid + it * LTHR;
      kreg[it] = xreg[it] = make_float4(0.f, 0.f, 0.f, 0.f);
      const int key = c * KC + i / V4;
      if (i < KC * V4 && key < L) {
        const long long r = row0 + key;
        const int col = hh * DH + (i % V4) * 4;
        kreg[it] = __ldg(reinterpret_cast<const float4*>(a.K + r * a.ldk + col));
        xreg[it] = __ldg(reinterpret_cast<const float4*>(X + r * ldx + col));
      }
    }
    if (want_marks && tid < KC) {
      const int key = c * KC + tid;
      mreg = key < L ? __ldg(reinterpret_cast<const uint4*>(a.marks + (row0 + key) * E)) : make_uint4(0u, 0u, 0u, 0u);
    }
  };
  // par = parity of the running chunk counter (the maxima are double-buffered); returns 1 / scale of K and X
  auto store_chunk = [&](int par, bool want_marks, float& isk, float& isx) {
    float mk = 0.f, mx = 0.f;
#pragma unroll
    for (int it = 0; it < KI; ++it) {
      mk = absmax4(mk, kreg[it]);
      mx = absmax4(mx, xreg[it]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mk = fmaxf(mk, __shfl_xor_sync(0xffffffffu, mk, o));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if (lane == 0) {
      atomicMax(&red[par * 4 + 0], __float_as_uint(mk));
      atomicMax(&red[par * 4 + 1], __float_as_uint(mx));
    }
    __syncthreads();  // (A) maxima complete; every warp has finished computing on the previous chunk's tiles
    float sk, sx;
    pow2_scale(__uint_as_float(red[par * 4 + 0]), sk, isk);
    pow2_scale(__uint_as_float(red[par * 4 + 1]), sx, isx);
    if (tid == 0) red[(par ^ 1) * 4 + 0] = red[(par ^ 1) * 4 + 1] = 0u;  // next chunk's cells (last read before (A))
#pragma unroll
    for (int it = 0; it < KI; ++it) {
      const int i = tid + it * LTHR;
      if (i < KC * V4) {  // keys >= L were loaded as zeros and are written as zeros
        const int k = i / V4, d0 = (i % V4) * 4;
        put4(Ks + k * RB, d0, DH * 2, kreg[it], sk);
        put4(Xs + k * RB, d0, DH * 2, xreg[it], sx);
      }
    }
    if (want_marks && tid < KC) {
      // marks (tf.to_float, temporal.py:311) as fp16 in B-fragment order; bytes -> fp16 exactly through
      // 1024 + b (0x6400 | b) - 1024  (see attn_f16.cu)
      const uint4 m = mreg;
      auto h2 = [](uint32_t w, uint32_t sel) {
        const uint32_t x = __byte_perm(w, 0x64646464u, sel);
        const __half2 r = __hsub2(*reinterpret_cast<const __half2*>(&x), __floats2half2_rn(1024.f, 1024.f));
        return *reinterpret_cast<const uint32_t*>(&r);
      };
      uint4 w0, w1;
      w0.x = h2(__byte_perm(m.x, m.y, 0x40u), 0x4140u); w0.y = h2(__byte_perm(m.z, m.w, 0x40u), 0x4140u);
      w0.z = h2(__byte_perm(m.x, m.y, 0x51u), 0x4140u); w0.w = h2(__byte_perm(m.z, m.w, 0x51u), 0x4140u);
      w1.x = h2(__byte_perm(m.x, m.y, 0x62u), 0x4140u); w1.y = h2(__byte_perm(m.z, m.w, 0x62u), 0x4140u);
      w1.z = h2(__byte_perm(m.x, m.y, 0x73u), 0x4140u); w1.w = h2(__byte_perm(m.z, m.w, 0x73u), 0x4140u);
      *reinterpret_cast<uint4*>(Ms + tid * 8) = w0;
      *reinterpret_cast<uint4*>(Ms + tid * 8 + 4) = w1;
    }
  };

  const float sc2 = kLog2e / sqrtf((float)DH);  // temporal.py:355,422; scores in the log2 domain
  // S chunk = Q K_c^T (accumulators P[nt][c]: rows g / g+8, keys nt*8 + 2t + (c&1) of the chunk), scaled and masked
  auto scores = [&](int c, float isk, float (&P)[NTC][4], float& cma, float& cmb) {
#pragma unroll
    for (int nt = 0; nt < NTC; ++nt) P[nt][0] = P[nt][1] = P[nt][2] = P[nt][3] = 0.f;
#pragma unroll
    for (int n0 = 0; n0 < NTC; n0 += 4) {
#pragma unroll
      for (int ks = 0; ks < KD; ++ks) {
        uint32_t kb[4][4];  // {b0 hi, b1 hi, b0 lo, b1 lo}
#pragma unroll
        for (int j = 0; j < 4; ++j) ldsm_x4(kb[j], ks_addr + (n0 + j) * 8 * RB + ks * 32);
#pragma unroll
        for (int j = 0; j < 4; ++j) mma_f16(P[n0 + j], ql[ks], kb[j][0], kb[j][1]);
#pragma unroll
        for (int j = 0; j < 4; ++j) mma_f16(P[n0 + j], qh[ks], kb[j][2], kb[j][3]);
#pragma unroll
        for (int j = 0; j < 4; ++j) mma_f16(P[n0 + j], qh[ks], kb[j][0], kb[j][1]);
      }
    }
    const f32x2 sca2 = pk2(sc2 * iqa * isk, sc2 * iqa * isk), scb2 = pk2(sc2 * iqb * isk, sc2 * iqb * isk);
    cma = cmb = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < NTC; ++nt) {
      const float2 kmv = *reinterpret_cast<const float2*>(km + c * KC + nt * 8 + 2 * t);
      upk2(mul2(pk2(P[nt][0], P[nt][1]), sca2), P[nt][0], P[nt][1]);
      upk2(mul2(pk2(P[nt][2], P[nt][3]), scb2), P[nt][2], P[nt][3]);
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        float s = fminf(P[nt][cc], (cc & 1) ? kmv.y : kmv.x);
        if (a.causal) {
          const int col = c * KC + nt * 8 + 2 * t + (cc & 1);
          if (col > ((cc < 2) ? qa : qb)) s = fminf(s, kFillMma);  // temporal.py:362-367
        }
        P[nt][cc] = s;
        if (cc < 2) cma = fmaxf(cma, s); else cmb = fmaxf(cmb, s);
      }
    }
    cma = fmaxf(cma, __shfl_xor_sync(0xffffffffu, cma, 1));
    cma = fmaxf(cma, __shfl_xor_sync(0xffffffffu, cma, 2));
    cmb = fmaxf(cmb, __shfl_xor_sync(0xffffffffu, cmb, 1));
    cmb = fmaxf(cmb, __shfl_xor_sync(0xffffffffu, cmb, 2));
  };

  // ---------------------------------------------------------------- pass 1: online softmax, H = P T
  int cc_run = 0;  // running chunk counter over both passes
  float m_a = -INFINITY, m_b = -INFINITY, l_a = 0.f, l_b = 0.f;  // l = 2^14 * (sum of 2^(s - m)), per-lane partial
  float H[ND][4];
#pragma unroll
  for (int n = 0; n < ND; ++n) H[n][0] = H[n][1] = H[n][2] = H[n][3] = 0.f;
  load_chunk(0, a.T, a.ldt, false);
  for (int c = 0; c < nchunks; ++c, ++cc_run) {
    float isk, ist;
    store_chunk(cc_run & 1, false, isk, ist);
    if (c + 1 < nchunks) load_chunk(c + 1, a.T, a.ldt, false);
    else load_chunk(0, a.V, a.ldv, true);      // first chunk of pass 2
    if (c == nchunks - 1) cp_async_wait_all();  // the MLP constants: visible to everybody after (B)
    __syncthreads();                            // (B) tiles written
    if (active) {
      float P[NTC][4], cma, cmb;
      scores(c, isk, P, cma, cmb);
      const float mna = fmaxf(m_a, cma), mnb = fmaxf(m_b, cmb);  // finite: key 0 exists, and masks are finite fills
      const float ala = ex2_approx(m_a - mna), alb = ex2_approx(m_b - mnb);
      m_a = mna; m_b = mnb;
      // P' = 2^14 2^(s - m): unnormalised probabilities in the fp16 range
      {
        f32x2 la2 = pk2(0.f, 0.f), lb2 = la2;
        const f32x2 oa2 = pk2(14.f - mna, 14.f - mna), ob2 = pk2(14.f - mnb, 14.f - mnb);
#pragma unroll
        for (int nt = 0; nt < NTC; ++nt) {
          float d0, d1, d2, d3;
          upk2(add2(pk2(P[nt][0], P[nt][1]), oa2), d0, d1);
          upk2(add2(pk2(P[nt][2], P[nt][3]), ob2), d2, d3);
          P[nt][0] = ex2_approx(d0); P[nt][1] = ex2_approx(d1);
          P[nt][2] = ex2_approx(d2); P[nt][3] = ex2_approx(d3);
          la2 = add2(la2, pk2(P[nt][0], P[nt][1]));
          lb2 = add2(lb2, pk2(P[nt][2], P[nt][3]));
        }
        float s0, s1;
        upk2(la2, s0, s1); l_a = fmaf(l_a, ala, s0 + s1);
        upk2(lb2, s0, s1); l_b = fmaf(l_b, alb, s0 + s1);
      }
      float Hc[ND][4];
      pv_product16<DH, NTC>(P, xs_addr, Hc);  // = 2^14 * scale(T_c) * (P_c T_c)
      const f32x2 al_a2 = pk2(ala, ala), al_b2 = pk2(alb, alb), ist2 = pk2(ist, ist);
#pragma unroll
      for (int n = 0; n < ND; ++n) {
        upk2(fma2(pk2(Hc[n][0], Hc[n][1]), ist2, mul2(pk2(H[n][0], H[n][1]), al_a2)), H[n][0], H[n][1]);
        upk2(fma2(pk2(Hc[n][2], Hc[n][3]), ist2, mul2(pk2(H[n][2], H[n][3]), al_b2)), H[n][2], H[n][3]);
      }
    }
  }

  // ---------------------------------------------------------------- intensity MLP -> lam (temporal.py:287-306)
  float linva = 0.f, linvb = 0.f;  // 1 / l  (l carries the factor 2^14)
  uint32_t lh[4], ll[4];
  float isla = 1.f, islb = 1.f, sla = 1.f, slb = 1.f;
  if (active) {
    l_a += __shfl_xor_sync(0xffffffffu, l_a, 1);
    l_a += __shfl_xor_sync(0xffffffffu, l_a, 2);
    l_b += __shfl_xor_sync(0xffffffffu, l_b, 1);
    l_b += __shfl_xor_sync(0xffffffffu, l_b, 2);
    linva = __frcp_rn(l_a);
    linvb = __frcp_rn(l_b);
    // H (true scale) and its row maxima: the A operand of the MLP gets a per-row power-of-two scale
    float hma = 0.f, hmb = 0.f;
#pragma unroll
    for (int n = 0; n < ND; ++n) {
      H[n][0] *= linva; H[n][1] *= linva; H[n][2] *= linvb; H[n][3] *= linvb;
      hma = fmaxf(hma, fmaxf(fabsf(H[n][0]), fabsf(H[n][1])));
      hmb = fmaxf(hmb, fmaxf(fabsf(H[n][2]), fabsf(H[n][3])));
    }
    hma = fmaxf(hma, __shfl_xor_sync(0xffffffffu, hma, 1));
    hma = fmaxf(hma, __shfl_xor_sync(0xffffffffu, hma, 2));
    hmb = fmaxf(hmb, __shfl_xor_sync(0xffffffffu, hmb, 1));
    hmb = fmaxf(hmb, __shfl_xor_sync(0xffffffffu, hmb, 2));
    float sha, isha, shb, ishb;
    pow2_scale(hma, sha, isha);
    pow2_scale(hmb, shb, ishb);
    uint32_t hh_[KD][4], hl_[KD][4];
#pragma unroll
    for (int ks = 0; ks < KD; ++ks) {
      split2(H[2 * ks][0] * sha, H[2 * ks][1] * sha, hh_[ks][0], hl_[ks][0]);
      split2(H[2 * ks][2] * shb, H[2 * ks][3] * shb, hh_[ks][1], hl_[ks][1]);
      split2(H[2 * ks + 1][0] * sha, H[2 * ks + 1][1] * sha, hh_[ks][2], hl_[ks][2]);
      split2(H[2 * ks + 1][2] * shb, H[2 * ks + 1][3] * shb, hh_[ks][3], hl_[ks][3]);
    }
    const float isw = misc[0];
    const f32x2 zsa2 = pk2(isha * isw, isha * isw), zsb2 = pk2(ishb * isw, ishb * isw);  // accumulator -> -z log2(e)
    // Events four at a time in a rolled loop; a 4x4 transpose-reduce over the quad leaves lane t with the sum of event
    // 4*eg + t, i.e. after the loop lane t owns the events t, 4+t, 8+t, 12+t = the k slots 2t, 2t+1, 2t+8, 2t+9 of its
    // A-fragment registers (the marks rows are staged in the same slot order).  Same code as attn_f16.cu.
    float va[4], vb[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) va[i] = vb[i] = 0.f;
#pragma unroll 1
    for (int eg = 0; eg < E / 4; ++eg) {
      f32x2 pa2[4], pb2[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) pa2[i] = pb2[i] = pk2(0.f, 0.f);
      const uint32_t* w1g = W1t + (size_t)(eg * 4 * ND * 8 + g) * SKW + t * 4;
      const float4* bwg = bw + eg * 4 * ND * 4 + t;
      const float* wvg = wv + eg * 4 * ND * 8 + 2 * t;
#pragma unroll
      for (int tq = 0; tq < ND; ++tq) {  // 4 tiles of 8 columns at a time
        float z[4][4];
#pragma unroll
        for (int j = 0; j < 4; ++j) z[j][0] = z[j][1] = z[j][2] = z[j][3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < KD; ++ks) {
          uint4 wb[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) wb[j] = *reinterpret_cast<const uint4*>(w1g + (tq * 4 + j) * 8 * SKW + ks * 16);
#pragma unroll
          for (int j = 0; j < 4; ++j) mma_f16(z[j], hl_[ks], wb[j].x, wb[j].y);
#pragma unroll
          for (int j = 0; j < 4; ++j) mma_f16(z[j], hh_[ks], wb[j].z, wb[j].w);
#pragma unroll
          for (int j = 0; j < 4; ++j) mma_f16(z[j], hh_[ks], wb[j].x, wb[j].y);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int tl = tq * 4 + j;      // tile inside the group; its event is tl / ND
          const float4 bb = bwg[tl * 4];  // {b1[c0], b1[c0+1], wsp[c0], wsp[c0+1]}, c0 = tile*8 + 2t
          const float2 we = *reinterpret_cast<const float2*>(wvg + tl * 8);
          const f32x2 b2 = pk2(bb.x, bb.y), w2 = pk2(bb.z, bb.w), we2 = pk2(we.x, we.y);
          float z0, z1, z2, z3;
          upk2(fma2(pk2(z[j][0], z[j][1]), zsa2, fma2(pk2(spa, spa), w2, b2)), z0, z1);
          upk2(fma2(pk2(z[j][2], z[j][3]), zsb2, fma2(pk2(spb, spb), w2, b2)), z2, z3);
          // z* hold -z*log2(e): sigmoid = 1 / (1 + 2^(z*))   (tf.nn.sigmoid, temporal.py:290)
          float a0, a1, a2, a3;
          upk2(add2(pk2(ex2_approx(z0), ex2_approx(z1)), pk2(1.f, 1.f)), a0, a1);
          upk2(add2(pk2(ex2_approx(z2), ex2_approx(z3)), pk2(1.f, 1.f)), a2, a3);
          pa2[tl / ND] = fma2(pk2(rcp_approx(a0), rcp_approx(a1)), we2, pa2[tl / ND]);
          pb2[tl / ND] = fma2(pk2(rcp_approx(a2), rcp_approx(a3)), we2, pb2[tl / ND]);
        }
      }
      float pa[4], pb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float s0, s1;
        upk2(pa2[i], s0, s1); pa[i] = s0 + s1;
        upk2(pb2[i], s0, s1); pb[i] = s0 + s1;
      }
      const bool odd = (t & 1) != 0, up = (t & 2) != 0;
      float xa, xb;
      {
        const float r0 = __shfl_xor_sync(0xffffffffu, odd ? pa[0] : pa[1], 1);
        const float r1 = __shfl_xor_sync(0xffffffffu, odd ? pa[2] : pa[3], 1);
        const float w0 = (odd ? pa[1] : pa[0]) + r0, w1 = (odd ? pa[3] : pa[2]) + r1;
        xa = (up ? w1 : w0) + __shfl_xor_sync(0xffffffffu, up ? w0 : w1, 2);
      }
      {
        const float r0 = __shfl_xor_sync(0xffffffffu, odd ? pb[0] : pb[1], 1);
        const float r1 = __shfl_xor_sync(0xffffffffu, odd ? pb[2] : pb[3], 1);
        const float w0 = (odd ? pb[1] : pb[0]) + r0, w1 = (odd ? pb[3] : pb[2]) + r1;
        xb = (up ? w1 : w0) + __shfl_xor_sync(0xffffffffu, up ? w0 : w1, 2);
      }
      const int ev = eg * 4 + t;
      const float s = sc[ev];
      // lam_e = s_e log(1 + exp(x / s_e)): the naive softplus of the reference (overflows to inf for x/s > 88.7, Q6)
      const float rs = rcp_approx(s) * kLog2e, sl = s * 0.69314718055994531f;
      const float la_ = sl * lg2_approx(1.f + ex2_approx(xa * rs));
      const float lb_ = sl * lg2_approx(1.f + ex2_approx(xb * rs));
      if (a.lam) {
        if (qa < L) a.lam[(((long long)hh * B + b) * L + qa) * E + ev] = la_;  // head-major, temporal.py:413
        if (qb < L) a.lam[(((long long)hh * B + b) * L + qb) * E + ev] = lb_;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        va[i] = (eg == i) ? la_ : va[i];
        vb[i] = (eg == i) ? lb_ : vb[i];
      }
    }
    float ma2 = fmaxf(fmaxf(fabsf(va[0]), fabsf(va[1])), fmaxf(fabsf(va[2]), fabsf(va[3])));
    float mb2 = fmaxf(fmaxf(fabsf(vb[0]), fabsf(vb[1])), fmaxf(fabsf(vb[2]), fabsf(vb[3])));
    ma2 = fmaxf(ma2, __shfl_xor_sync(0xffffffffu, ma2, 1));
    ma2 = fmaxf(ma2, __shfl_xor_sync(0xffffffffu, ma2, 2));
    mb2 = fmaxf(mb2, __shfl_xor_sync(0xffffffffu, mb2, 1));
    mb2 = fmaxf(mb2, __shfl_xor_sync(0xffffffffu, mb2, 2));
    pow2_scale(ma2, sla, isla);
    pow2_scale(mb2, slb, islb);
    split2(va[0] * sla, va[1] * sla, lh[0], ll[0]);  // a0: row g,   k slots 2t, 2t+1   = events t, 4+t
    split2(vb[0] * slb, vb[1] * slb, lh[1], ll[1]);  // a1: row g+8
    split2(va[2] * sla, va[3] * sla, lh[2], ll[2]);  // a2: row g,   k slots 2t+8, 2t+9 = events 8+t, 12+t
    split2(vb[2] * slb, vb[3] * slb, lh[3], ll[3]);  // a3: row g+8
  }

  // ---------------------------------------------------------------- pass 2: O = (G o P) V
  float O[ND][4];
#pragma unroll
  for (int n = 0; n < ND; ++n) O[n][0] = O[n][1] = O[n][2] = O[n][3] = 0.f;
  const f32x2 pna2 = pk2(16384.f * linva, 16384.f * linva), pnb2 = pk2(16384.f * linvb, 16384.f * linvb);
  for (int c = 0; c < nchunks; ++c, ++cc_run) {
    float isk, isv;
    store_chunk(cc_run & 1, true, isk, isv);
    if (c + 1 < nchunks) load_chunk(c + 1, a.V, a.ldv, true);
    __syncthreads();  // (B)
    if (active) {
      float P[NTC][4], cma, cmb;
      scores(c, isk, P, cma, cmb);
      // P' = 2^14 2^(s - m) / sum: <= 2^14
      {
        const f32x2 oa2 = pk2(14.f - m_a, 14.f - m_a), ob2 = pk2(14.f - m_b, 14.f - m_b);
#pragma unroll
        for (int nt = 0; nt < NTC; ++nt) {
          float d0, d1, d2, d3;
          upk2(add2(pk2(P[nt][0], P[nt][1]), oa2), d0, d1);
          upk2(add2(pk2(P[nt][2], P[nt][3]), ob2), d2, d3);
          upk2(mul2(pk2(ex2_approx(d0), ex2_approx(d1)), pna2), P[nt][0], P[nt][1]);
          upk2(mul2(pk2(ex2_approx(d2), ex2_approx(d3)), pnb2), P[nt][2], P[nt][3]);
        }
      }
      // G_c = lam M_c^T (marks exact in fp16: 2 MMAs), set_diag, gate  (temporal.py:309-313, 438-441); the
      // accumulators carry the row scale of lam, so a forced diagonal of 1 is that scale
      float ga = 0.f, gb = 0.f;  // chunk row maxima of G o P
#pragma unroll
      for (int n0 = 0; n0 < NTC; n0 += 4) {
        float G[4][4];
#pragma unroll
        for (int j = 0; j < 4; ++j) G[j][0] = G[j][1] = G[j][2] = G[j][3] = 0.f;
        uint2 mk[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) mk[j] = *reinterpret_cast<const uint2*>(Ms + ((n0 + j) * 8 + g) * 8 + 2 * t);
#pragma unroll
        for (int j = 0; j < 4; ++j) mma_f16(G[j], ll, mk[j].x, mk[j].y);
#pragma unroll
        for (int j = 0; j < 4; ++j) mma_f16(G[j], lh, mk[j].x, mk[j].y);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (a.diag_one) {
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
              const int col = c * KC + (n0 + j) * 8 + 2 * t + (cc & 1);
              G[j][cc] = (col == ((cc < 2) ? qa : qb)) ? ((cc < 2) ? sla : slb) : G[j][cc];
            }
          }
          upk2(mul2(pk2(P[n0 + j][0], P[n0 + j][1]), pk2(G[j][0], G[j][1])), P[n0 + j][0], P[n0 + j][1]);
          upk2(mul2(pk2(P[n0 + j][2], P[n0 + j][3]), pk2(G[j][2], G[j][3])), P[n0 + j][2], P[n0 + j][3]);
          ga = fmaxf(ga, fmaxf(fabsf(P[n0 + j][0]), fabsf(P[n0 + j][1])));
          gb = fmaxf(gb, fmaxf(fabsf(P[n0 + j][2]), fabsf(P[n0 + j][3])));
        }
      }
      ga = fmaxf(ga, __shfl_xor_sync(0xffffffffu, ga, 1));
      ga = fmaxf(ga, __shfl_xor_sync(0xffffffffu, ga, 2));
      gb = fmaxf(gb, __shfl_xor_sync(0xffffffffu, gb, 1));
      gb = fmaxf(gb, __shfl_xor_sync(0xffffffffu, gb, 2));
      float sga, isga, sgb, isgb;
      pow2_scale(ga, sga, isga);
      pow2_scale(gb, sgb, isgb);
#pragma unroll
      for (int nt = 0; nt < NTC; ++nt) {
        upk2(mul2(pk2(P[nt][0], P[nt][1]), pk2(sga, sga)), P[nt][0], P[nt][1]);
        upk2(mul2(pk2(P[nt][2], P[nt][3]), pk2(sgb, sgb)), P[nt][2], P[nt][3]);
      }
      float Oc[ND][4];
      pv_product16<DH, NTC>(P, xs_addr, Oc);  // = 2^14 scale(lam row) scale(G o P chunk row) scale(V_c) * O_c
      const f32x2 fa2 = pk2(isv * isga, isv * isga), fb2 = pk2(isv * isgb, isv * isgb);
#pragma unroll
      for (int n = 0; n < ND; ++n) {
        upk2(fma2(pk2(Oc[n][0], Oc[n][1]), fa2, pk2(O[n][0], O[n][1])), O[n][0], O[n][1]);
        upk2(fma2(pk2(Oc[n][2], Oc[n][3]), fb2, pk2(O[n][2], O[n][3])), O[n][2], O[n][3]);
      }
    }
  }

  // ---------------------------------------------------------------- residual + store (temporal.py:385,447)
  if (active) {
    constexpr float k2m14 = 1.0f / 16384.f;
    const float fa = k2m14 * isla, fb = k2m14 * islb;
    float omax = 0.f;
#pragma unroll
    for (int n = 0; n < ND; ++n) {
      const int col = hh * DH + n * 8 + 2 * t;
      if (qa < L) {
        float2 o = make_float2(O[n][0] * fa, O[n][1] * fa);
        if (a.R) {
          const float2 r = __ldg(reinterpret_cast<const float2*>(a.R + ra * a.ldr + col));
          o.x += r.x; o.y += r.y;
        }
        omax = fmaxf(omax, fmaxf(fabsf(o.x), fabsf(o.y)));
        *reinterpret_cast<float2*>(a.O + (row0 + qa) * a.ldo + col) = o;
      }
      if (qb < L) {
        float2 o = make_float2(O[n][2] * fb, O[n][3] * fb);
        if (a.R) {
          const float2 r = __ldg(reinterpret_cast<const float2*>(a.R + rb * a.ldr + col));
          o.x += r.x; o.y += r.y;
        }
        omax = fmaxf(omax, fmaxf(fabsf(o.x), fabsf(o.y)));
        *reinterpret_cast<float2*>(a.O + (row0 + qb) * a.ldo + col) = o;
      }
    }
    if (a.out_amax) amax_publish(a.out_amax, omax, lane);  // consumed by the scaled 3xFP16 attention-out GEMM
  }
}

template <int DH>
int launch_long_t(const AttnArgs& a, cudaStream_t st) {
  using LL = LongLayout<DH>;
  const int nchunks = (a.L + KC - 1) / KC;
  const int qblocks = (a.L + QROWS - 1) / QROWS;
  const size_t smem = LL::bytes(nchunks * KC);
  if (smem > 227 * 1024) return 1;
  auto kern = attention_f16_long_kernel<DH>;
  EDGL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<(unsigned)((long long)a.B * a.h * qblocks), LTHR, smem, st>>>(a, qblocks, nchunks);
  EDGL_LAUNCH_CHECK();
  return 0;
}

}  // namespace

// returns 0 = launched, 1 = shape not covered, <0 = error
int launch_attention_f16_long(const AttnArgs& a, cudaStream_t st) {
  const int dh = a.d / a.h;
  if ((dh != 16 && dh != 32) || a.E != 16 || !a.mlp_pack || a.L < 1) return 1;
  if (reinterpret_cast<uintptr_t>(a.marks) & 15) return 1;  // mark rows are read as 16-byte words
  if ((long long)a.B * a.h * ((a.L + QROWS - 1) / QROWS) > 0x7fffffffll) return 1;
  return dh == 16 ? launch_long_t<16>(a, st) : launch_long_t<32>(a, st);
}

}  // namespace edgl
