// attn_f16.cu - the fused self-modulating attention core (temporal.py:345-385 / 412-447 / 281-315) on
// mma.sync.m16n8k16 with a SCALED 3xFP16 split.  Same pipeline as attn_mma.cuh
//   S = Q K^T, P = softmax, H = P T, Z = sigmoid([H,span] W1 + b1), lam = softplus, G = lam M^T, O = (G o P) V
// but every operand x is first multiplied by a power of two that puts the operand's maximum into
// [2^14, 2^15) and then written as x = hi + lo with hi = the top 11 significant bits (exact in fp16) and
// lo = x - hi rounded to fp16.  D += A_lo*B_hi + A_hi*B_lo + A_hi*B_hi in fp32 then has the accuracy of the
// 3xTF32 scheme (|error| <= max(2^-21 |x|, 2^-25 after scaling) per element; fp16 subnormals give the
// absolute floor) at HALF the tensor instructions (k = 16 per MMA instead of 8), half the shared-memory
// operand bytes and half the fragment loads.  The scales are exact powers of two and are divided out of the
// fp32 accumulators, so no rounding is added.  Where the scales come from:
//   Q, lam, G o P : per query row (A operands: a row scale is a row scale of the output), from the row max
//   K, V, T       : per (sequence, head) tile, from a block-wide max taken while staging
//   P             : 2^14 (P <= 1), folded into the softmax normaliser;  H : the scale of T (|H| <= max|T|)
//   W1            : per tensor, in a pack built once per edgl_commit (mlp_pack_kernel)
// Only dh = 16*k, E = 16 is instantiated; other shapes keep the TF32 kernel (attn_mma.cuh).
#include "attn_f16_common.cuh"

namespace edgl {
namespace {

using namespace f16c;

// NT = number of 8-key tiles held in registers (L <= 8*NT); HPC heads are processed by one CTA in turn.
// MAXREG bounds the registers per thread and so the CTAs per SM (7 warps/CTA at L = 100: 128 -> 2, 80 -> 3, 72 -> 4).
// ABL: precision ablation (common.cuh; tools/ablation.py) - individual lo products left out.  0 on the product path.
template <int DH, int NT, int MAXREG, int ABL = 0>
__global__ void __maxnreg__(MAXREG) attention_f16_kernel(AttnArgs a, int hpc) {
  using LY = F16Layout<DH>;
  constexpr int E = LY::E, KD = LY::KD, ND = LY::ND, NC = LY::NC, MT = LY::MT, SKW = LY::SKW, RB = LY::RB;
  constexpr int NB = (NT + 1) / 2, LP = NT * 8, RP = NB * 16;
  constexpr int V4 = DH / 4;
  constexpr int KI = (NT * 8 * V4 + 255) / 256 < 2 ? 2 : (NT * 8 * V4 + 255) / 256;  // (key, 4 dims) items per thread
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const uint32_t* W1t = reinterpret_cast<const uint32_t*>(smem_raw);                  // [NC][SKW]
  const float4* bw = reinterpret_cast<const float4*>(smem_raw + LY::BW_OFF);          // [NC/2]
  const float* wv = reinterpret_cast<const float*>(smem_raw + LY::WV_OFF);            // [NC]
  const float* sc = reinterpret_cast<const float*>(smem_raw + LY::SC_OFF);            // [E]
  const float* misc = reinterpret_cast<const float*>(smem_raw + LY::MISC_OFF);
  unsigned char* Ks = smem_raw + LY::PACK_BYTES;                                      // [RP] rows of RB bytes
  unsigned char* Vs = Ks + RP * RB;
  unsigned char* Ts = Vs + RP * RB;
  uint32_t* Ms = reinterpret_cast<uint32_t*>(Ts + RP * RB);                           // [LP][8] marks as fp16
  float* km = reinterpret_cast<float*>(Ms + LP * 8);                                  // [LP] min-mask
  unsigned int* red = reinterpret_cast<unsigned int*>(km + LP);                       // [4] tile maxima (K, V, T)

  const int L = a.L, B = a.B;
  const int groups = a.h / hpc;
  const int b = blockIdx.x / groups, hh0 = (blockIdx.x % groups) * hpc;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const long long row0 = (long long)b * L;

  // ---------------------------------------------------------------- per-sequence operands
  // MLP constants: asynchronous global -> shared copy, overlapped with everything up to the first barrier wait
  {
    const unsigned char* src = reinterpret_cast<const unsigned char*>(a.mlp_pack);
    for (int i = tid; i < LY::PACK_BYTES / 16; i += nthr) cp_async16(smem_raw + i * 16, src + i * 16);
  }
  // marks (tf.to_float, temporal.py:311) as fp16; a row holds the two B-fragment words of lane t = 0..3 back to back,
  // so one 8-byte load is the fragment.  Bytes -> fp16 exactly through 1024 + b (0x6400 | b) - 1024.
  for (int k = tid; k < LP; k += nthr) {
    uint4 m = make_uint4(0u, 0u, 0u, 0u);
    float kv = -INFINITY;
    if (k < L) {
      m = __ldg(reinterpret_cast<const uint4*>(a.marks + (row0 + k) * E));
      kv = a.kmask[row0 + k] ? INFINITY : kFillMma;
    }
    auto h2 = [](uint32_t w, uint32_t sel) {
      const uint32_t x = __byte_perm(w, 0x64646464u, sel);
      const __half2 r = __hsub2(*reinterpret_cast<const __half2*>(&x), __floats2half2_rn(1024.f, 1024.f));
      return *reinterpret_cast<const uint32_t*>(&r);
    };
    // word 2t = events (t, 4+t), word 2t+1 = events (8+t, 12+t): k slots 2t, 2t+1, 2t+8, 2t+9 of lane t (see the MLP loop)
    uint4 w0, w1;
    w0.x = h2(__byte_perm(m.x, m.y, 0x40u), 0x4140u); w0.y = h2(__byte_perm(m.z, m.w, 0x40u), 0x4140u);
    w0.z = h2(__byte_perm(m.x, m.y, 0x51u), 0x4140u); w0.w = h2(__byte_perm(m.z, m.w, 0x51u), 0x4140u);
    w1.x = h2(__byte_perm(m.x, m.y, 0x62u), 0x4140u); w1.y = h2(__byte_perm(m.z, m.w, 0x62u), 0x4140u);
    w1.z = h2(__byte_perm(m.x, m.y, 0x73u), 0x4140u); w1.w = h2(__byte_perm(m.z, m.w, 0x73u), 0x4140u);
    *reinterpret_cast<uint4*>(Ms + k * 8) = w0;
    *reinterpret_cast<uint4*>(Ms + k * 8 + 4) = w1;
    km[k] = kv;  // min-mask: +inf real key, fill = masked id, -inf = beyond L
  }
  // rows L..RP-1 of K / V / T stay zero for every head
  for (int i = tid; i < (RP - L) * (RB / 16) * 3; i += nthr) {
    const int which = i / ((RP - L) * (RB / 16)), rem = i % ((RP - L) * (RB / 16));
    *reinterpret_cast<uint4*>(Ks + which * RP * RB + (L + rem / (RB / 16)) * RB + (rem % (RB / 16)) * 16) =
        make_uint4(0u, 0u, 0u, 0u);
  }

  const float inv_sqrt_dh = 1.0f / sqrtf((float)DH);  // temporal.py:355,422
  const int num_mt = (L + 15) >> 4;
  // this lane's ldmatrix row addresses (matrix m = lane / 8, row r = lane % 8)
  const int lm = lane >> 3, lr = lane & 7;
  const uint32_t ks_addr = (uint32_t)__cvta_generic_to_shared(Ks) + lr * RB + (lm & 1) * 16 + (lm >> 1) * (DH * 2);
  const uint32_t xoff = ((lm & 1) * 8 + lr) * RB + (lm >> 1) * 16;
  const uint32_t vs_addr = (uint32_t)__cvta_generic_to_shared(Vs) + xoff;
  const uint32_t ts_addr = (uint32_t)__cvta_generic_to_shared(Ts) + xoff;

  for (int hh = hh0; hh < hh0 + hpc; ++hh) {
    // ---------------------------------------------------------------- stage K, V, T of this head
    // every global load of the phase is issued before the first use: K/V/T slices, and the Q rows of this warp's
    // first tile
    float4 kreg[KI], vreg[KI], treg[KI];
    float2 xa[KD][2], xb[KD][2];  // raw Q values of rows g / g+8 of a tile (k slots 2t,2t+1 / 2t+8,2t+9 per k16 step)
    float spa = 0.f, spb = 0.f;   // intervals of the two rows
    auto load_q = [&](int mt_) {
      const int qa_ = mt_ * 16 + g, qb_ = qa_ + 8;
      const long long ra_ = row0 + (qa_ < L ? qa_ : L - 1), rb_ = row0 + (qb_ < L ? qb_ : L - 1);
#pragma unroll
      for (int ks = 0; ks < KD; ++ks)
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          xa[ks][hf] = __ldg(reinterpret_cast<const float2*>(a.Q + ra_ * a.ldq + hh * DH + ks * 16 + hf * 8 + 2 * t));
          xb[ks][hf] = __ldg(reinterpret_cast<const float2*>(a.Q + rb_ * a.ldq + hh * DH + ks * 16 + hf * 8 + 2 * t));
        }
      spa = __ldg(a.spans + ra_);
      spb = __ldg(a.spans + rb_);
    };
    if (warp < num_mt) load_q(warp);
    float mk = 0.f, mv = 0.f, mtt = 0.f;
#pragma unroll
    for (int it = 0; it < KI; ++it) {
      const int i = tid + it * nthr;
      kreg[it] = vreg[it] = treg[it] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < L * V4) {
        const long long r = row0 + i / V4;
        const int col = hh * DH + (i % V4) * 4;
        kreg[it] = __ldg(reinterpret_cast<const float4*>(a.K + r * a.ldk + col));
        vreg[it] = __ldg(reinterpret_cast<const float4*>(a.V + r * a.ldv + col));
        treg[it] = __ldg(reinterpret_cast<const float4*>(a.T + r * a.ldt + col));
      }
    }
    if (tid < 4) red[tid] = 0u;
    __syncthreads();  // previous head fully consumed; red zeroed
#pragma unroll
    for (int it = 0; it < KI; ++it) {
      mk = absmax4(mk, kreg[it]);
      mv = absmax4(mv, vreg[it]);
      mtt = absmax4(mtt, treg[it]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mk = fmaxf(mk, __shfl_xor_sync(0xffffffffu, mk, o));
      mv = fmaxf(mv, __shfl_xor_sync(0xffffffffu, mv, o));
      mtt = fmaxf(mtt, __shfl_xor_sync(0xffffffffu, mtt, o));
    }
    if (lane == 0) {
      atomicMax(&red[0], __float_as_uint(mk));
      atomicMax(&red[1], __float_as_uint(mv));
      atomicMax(&red[2], __float_as_uint(mtt));
    }
    __syncthreads();
    float sk, isk, sv, isv, st, ist;
    pow2_scale(__uint_as_float(red[0]), sk, isk);
    pow2_scale(__uint_as_float(red[1]), sv, isv);
    pow2_scale(__uint_as_float(red[2]), st, ist);
    // rows [key]{hi[DH] | lo[DH] | pad}: natural dim order, read with ldmatrix (K) / ldmatrix.trans (V, T)
#pragma unroll
    for (int it = 0; it < KI; ++it) {
      const int i = tid + it * nthr;
      if (i < L * V4) {
        const int k = i / V4, d0 = (i % V4) * 4;
        put4(Ks + k * RB, d0, DH * 2, kreg[it], sk);
        put4(Vs + k * RB, d0, DH * 2, vreg[it], sv);
        put4(Ts + k * RB, d0, DH * 2, treg[it], st);
      }
    }
    cp_async_wait_all();
    __syncthreads();
    const float isw = misc[0];
    const float zscale = ist * isw;  // MLP accumulator -> -z log2(e)

    for (int mt = warp; mt < num_mt; mt += (nthr >> 5)) {
      const int q0 = mt * 16;
      const int qa = q0 + g, qb = q0 + g + 8;  // this thread's two query rows
      const long long ra = row0 + (qa < L ? qa : L - 1), rb = row0 + (qb < L ? qb : L - 1);

      // ---- Q fragments (A operand of S = Q K^T), scaled per row
      float iqa, iqb;
      uint32_t qh[KD][4], ql[KD][4];
      {
        if (mt != warp) load_q(mt);  // the first tile's rows were requested before the staging barriers
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
      // ---- S = Q K^T  (accumulators P[nt][c]: rows g / g+8, keys nt*8 + 2t + (c&1)); four key tiles in flight
      float P[NT][4];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) P[nt][0] = P[nt][1] = P[nt][2] = P[nt][3] = 0.f;
#pragma unroll
      for (int n0 = 0; n0 < NT; n0 += 4) {
#pragma unroll
        for (int ks = 0; ks < KD; ++ks) {
          uint32_t kb[4][4];  // {b0 hi, b1 hi, b0 lo, b1 lo}
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (n0 + j < NT) ldsm_x4(kb[j], ks_addr + (n0 + j) * 8 * RB + ks * 32);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if ((ABL & 1) == 0 && n0 + j < NT) mma_f16(P[n0 + j], ql[ks], kb[j][0], kb[j][1]);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if ((ABL & 2) == 0 && n0 + j < NT) mma_f16(P[n0 + j], qh[ks], kb[j][2], kb[j][3]);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (n0 + j < NT) mma_f16(P[n0 + j], qh[ks], kb[j][0], kb[j][1]);
        }
      }
      // ---- scale, key mask, causal mask, softmax in the log2 domain (see attn_mma.cuh)
      const float sc2 = inv_sqrt_dh * kLog2e * isk;
      const float sca = sc2 * iqa, scb = sc2 * iqb;
      float ma = -INFINITY, mb = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const float2 kmv = *reinterpret_cast<const float2*>(km + nt * 8 + 2 * t);
        upk2(mul2(pk2(P[nt][0], P[nt][1]), pk2(sca, sca)), P[nt][0], P[nt][1]);
        upk2(mul2(pk2(P[nt][2], P[nt][3]), pk2(scb, scb)), P[nt][2], P[nt][3]);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float s = fminf(P[nt][c], (c & 1) ? kmv.y : kmv.x);
          if (a.causal) {
            const int col = nt * 8 + 2 * t + (c & 1);
            if (col > ((c < 2) ? qa : qb)) s = fminf(s, kFillMma);  // temporal.py:362-367
          }
          P[nt][c] = s;
          if (c < 2) ma = fmaxf(ma, s); else mb = fmaxf(mb, s);
        }
      }
      ma = fmaxf(ma, __shfl_xor_sync(0xffffffffu, ma, 1));
      ma = fmaxf(ma, __shfl_xor_sync(0xffffffffu, ma, 2));
      mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, 1));
      mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, 2));
      float la, lb;
      {
        f32x2 la2 = pk2(0.f, 0.f), lb2 = la2;
        const f32x2 ma2 = pk2(ma, ma), mb2 = pk2(mb, mb);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          float d0, d1, d2, d3;
          upk2(sub2(pk2(P[nt][0], P[nt][1]), ma2), d0, d1);
          upk2(sub2(pk2(P[nt][2], P[nt][3]), mb2), d2, d3);
          P[nt][0] = ex2_approx(d0); P[nt][1] = ex2_approx(d1);
          P[nt][2] = ex2_approx(d2); P[nt][3] = ex2_approx(d3);
          la2 = add2(la2, pk2(P[nt][0], P[nt][1]));
          lb2 = add2(lb2, pk2(P[nt][2], P[nt][3]));
        }
        float l0, l1;
        upk2(la2, l0, l1); la = l0 + l1;
        upk2(lb2, l0, l1); lb = l0 + l1;
      }
      la += __shfl_xor_sync(0xffffffffu, la, 1);
      la += __shfl_xor_sync(0xffffffffu, la, 2);
      lb += __shfl_xor_sync(0xffffffffu, lb, 1);
      lb += __shfl_xor_sync(0xffffffffu, lb, 2);
      // P' = P * 2^14 (exact scaling of the rounded quotient p * (1/l))
      const float ia = __frcp_rn(la) * 16384.f, ib = __frcp_rn(lb) * 16384.f;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        upk2(mul2(pk2(P[nt][0], P[nt][1]), pk2(ia, ia)), P[nt][0], P[nt][1]);
        upk2(mul2(pk2(P[nt][2], P[nt][3]), pk2(ib, ib)), P[nt][2], P[nt][3]);
      }
      // ---- H = P T : accumulator = H * 2^14 * scale(T); |H * scale(T)| < 2^15 because rows of P sum to 1
      float H[ND][4];
      pv_product16<DH, NT, ABL>(P, ts_addr, H);
      // ---- intensity MLP: Z = sigmoid([H, span] W1 + b1); dot with w per event (temporal.py:287-305)
      uint32_t hh_[KD][4], hl_[KD][4];
#pragma unroll
      for (int ks = 0; ks < KD; ++ks) {
        constexpr float k2m14 = 1.0f / 16384.f;
        split2(H[2 * ks][0] * k2m14, H[2 * ks][1] * k2m14, hh_[ks][0], hl_[ks][0]);
        split2(H[2 * ks][2] * k2m14, H[2 * ks][3] * k2m14, hh_[ks][1], hl_[ks][1]);
        split2(H[2 * ks + 1][0] * k2m14, H[2 * ks + 1][1] * k2m14, hh_[ks][2], hl_[ks][2]);
        split2(H[2 * ks + 1][2] * k2m14, H[2 * ks + 1][3] * k2m14, hh_[ks][3], hl_[ks][3]);
      }
      // Events are processed four at a time in a ROLLED loop (code size: the instruction cache is shared by 28 warps in
      // different phases).  Within a group each lane accumulates its columns' share of the four per-event dot
      // products; a 4x4 transpose-reduce over the quad (3 shuffles per row) leaves lane t with the full sum of event
      // 4*eg + t, which it turns into lam right away.  So after the loop lane t owns the events t, 4+t, 8+t, 12+t:
      // these ARE the k slots 2t, 2t+1, 2t+8, 2t+9 of its A-fragment registers (the marks rows are staged in the same
      // slot order), and no per-event array is ever live.
      // lam_e = s_e log(1 + exp(x / s_e))  (temporal.py:305-306)
      float va[4], vb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) va[i] = vb[i] = 0.f;
#pragma unroll 1
      for (int eg = 0; eg < E / 4; ++eg) {
        f32x2 pa2[4], pb2[4];  // two-lane partial sums (even / odd columns) of the four events, rows a / b
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
            for (int j = 0; j < 4; ++j)
              wb[j] = *reinterpret_cast<const uint4*>(w1g + (tq * 4 + j) * 8 * SKW + ks * 16);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if ((ABL & 16) == 0) mma_f16(z[j], hl_[ks], wb[j].x, wb[j].y);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if ((ABL & 32) == 0) mma_f16(z[j], hh_[ks], wb[j].z, wb[j].w);
#pragma unroll
            for (int j = 0; j < 4; ++j) mma_f16(z[j], hh_[ks], wb[j].x, wb[j].y);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int tl = tq * 4 + j;         // tile inside the group; its event is tl / ND
            const float4 bb = bwg[tl * 4];     // {b1[c0], b1[c0+1], wsp[c0], wsp[c0+1]}, c0 = tile*8 + 2t
            const float2 we = *reinterpret_cast<const float2*>(wvg + tl * 8);
            const f32x2 b2 = pk2(bb.x, bb.y), w2 = pk2(bb.z, bb.w), zs2 = pk2(zscale, zscale), we2 = pk2(we.x, we.y);
            float z0, z1, z2, z3;
            upk2(fma2(pk2(z[j][0], z[j][1]), zs2, fma2(pk2(spa, spa), w2, b2)), z0, z1);
            upk2(fma2(pk2(z[j][2], z[j][3]), zs2, fma2(pk2(spb, spb), w2, b2)), z2, z3);
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
          float l0, l1;
          upk2(pa2[i], l0, l1); pa[i] = l0 + l1;
          upk2(pb2[i], l0, l1); pb[i] = l0 + l1;
        }
        // 4x4 transpose-reduce over the quad: lane t ends with the sum over lanes of p[t]
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
        // naive softplus like the reference (overflows to inf for x/s > 88.7, Q6), on the fast exp2/log2 units
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
      uint32_t lh[4], ll[4];
      float isla, islb, sla, slb;
      {
        float ma2 = fmaxf(fmaxf(fabsf(va[0]), fabsf(va[1])), fmaxf(fabsf(va[2]), fabsf(va[3])));
        float mb2 = fmaxf(fmaxf(fabsf(vb[0]), fabsf(vb[1])), fmaxf(fabsf(vb[2]), fabsf(vb[3])));
        ma2 = fmaxf(ma2, __shfl_xor_sync(0xffffffffu, ma2, 1));
        ma2 = fmaxf(ma2, __shfl_xor_sync(0xffffffffu, ma2, 2));
        mb2 = fmaxf(mb2, __shfl_xor_sync(0xffffffffu, mb2, 1));
        mb2 = fmaxf(mb2, __shfl_xor_sync(0xffffffffu, mb2, 2));
        pow2_scale(ma2, sla, isla);
        pow2_scale(mb2, slb, islb);
        split2(va[0] * sla, va[1] * sla, lh[0], ll[0]);   // a0: row g,   k slots 2t, 2t+1   = events t, 4+t
        split2(vb[0] * slb, vb[1] * slb, lh[1], ll[1]);   // a1: row g+8
        split2(va[2] * sla, va[3] * sla, lh[2], ll[2]);   // a2: row g,   k slots 2t+8, 2t+9 = events 8+t, 12+t
        split2(vb[2] * slb, vb[3] * slb, lh[3], ll[3]);   // a3: row g+8
      }
      // ---- G = lam M^T (marks exact in fp16: 2 MMAs), set_diag, gate: P <- G o P   (temporal.py:309-313,438-441)
      // accumulators carry the row scale of lam, so a forced diagonal of 1 is that scale
      float2 res_a[ND], res_b[ND];  // residual rows, requested here so the gate and (G o P) V hide the latency
#pragma unroll
      for (int n = 0; n < ND; ++n) {
        res_a[n] = res_b[n] = make_float2(0.f, 0.f);
        if (a.R) {
          res_a[n] = __ldg(reinterpret_cast<const float2*>(a.R + ra * a.ldr + hh * DH + n * 8 + 2 * t));
          res_b[n] = __ldg(reinterpret_cast<const float2*>(a.R + rb * a.ldr + hh * DH + n * 8 + 2 * t));
        }
      }
      float ga = 0.f, gb = 0.f;  // row maxima of G o P
#pragma unroll
      for (int n0 = 0; n0 < NT; n0 += 4) {
        float G[4][4];
#pragma unroll
        for (int j = 0; j < 4; ++j) G[j][0] = G[j][1] = G[j][2] = G[j][3] = 0.f;
        uint2 mk[4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (n0 + j < NT) mk[j] = *reinterpret_cast<const uint2*>(Ms + ((n0 + j) * 8 + g) * 8 + 2 * t);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if ((ABL & 64) == 0 && n0 + j < NT) mma_f16(G[j], ll, mk[j].x, mk[j].y);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (n0 + j < NT) mma_f16(G[j], lh, mk[j].x, mk[j].y);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (n0 + j < NT) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const int col = (n0 + j) * 8 + 2 * t + (c & 1);
              const int qr = (c < 2) ? qa : qb;
              G[j][c] = (a.diag_one && col == qr) ? ((c < 2) ? sla : slb) : G[j][c];
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
      for (int nt = 0; nt < NT; ++nt) {
        upk2(mul2(pk2(P[nt][0], P[nt][1]), pk2(sga, sga)), P[nt][0], P[nt][1]);
        upk2(mul2(pk2(P[nt][2], P[nt][3]), pk2(sgb, sgb)), P[nt][2], P[nt][3]);
      }
      // ---- O = (G o P) V : accumulator = O * 2^14 * scale(lam row) * scale(G o P row) * scale(V)
      float O[ND][4];
      pv_product16<DH, NT, ABL>(P, vs_addr, O);
      constexpr float k2m14 = 1.0f / 16384.f;
      const float fa = isv * isga, fb = isv * isgb;
      const float fa2 = k2m14 * isla, fb2 = k2m14 * islb;
      // ---- residual + store (temporal.py:385,447)
      float omax = 0.f;
#pragma unroll
      for (int n = 0; n < ND; ++n) {
        const int col = hh * DH + n * 8 + 2 * t;
        if (qa < L) {
          float2 o = make_float2(O[n][0] * fa * fa2, O[n][1] * fa * fa2);
          o.x += res_a[n].x; o.y += res_a[n].y;
          omax = fmaxf(omax, fmaxf(fabsf(o.x), fabsf(o.y)));
          *reinterpret_cast<float2*>(a.O + (row0 + qa) * a.ldo + col) = o;
        }
        if (qb < L) {
          float2 o = make_float2(O[n][2] * fb * fb2, O[n][3] * fb * fb2);
          o.x += res_b[n].x; o.y += res_b[n].y;
          omax = fmaxf(omax, fmaxf(fabsf(o.x), fabsf(o.y)));
          *reinterpret_cast<float2*>(a.O + (row0 + qb) * a.ldo + col) = o;
        }
      }
      if (a.out_amax) amax_publish(a.out_amax, omax, lane);  // consumed by the scaled 3xFP16 attention-out GEMM
    }
  }
}

template <int DH, int NT, int MAXREG, int ABL = 0>
int launch_f16_t(const AttnArgs& a, int hpc, cudaStream_t st) {
  using LY = F16Layout<DH>;
  const size_t smem = LY::smem_bytes(NT);
  if (smem > 227 * 1024) return 1;
  auto kern = attention_f16_kernel<DH, NT, MAXREG, ABL>;
  EDGL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // several ~50 KB CTAs per SM: ask for the largest shared-memory carve-out, or the driver's default split caps residency
  EDGL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  int warps = (a.L + 15) / 16;
  if (warps > 8) warps = 8;
  kern<<<(unsigned)(a.B * (a.h / hpc)), warps * 32, smem, st>>>(a, hpc);
  EDGL_LAUNCH_CHECK();
  return 0;
}

}  // namespace

size_t attention_f16_pack_bytes(int dh, int E) {
  if (E != 16) return 0;
  if (dh == 16) return F16Layout<16>::PACK_BYTES;
  if (dh == 32) return F16Layout<32>::PACK_BYTES;  // attn_f16_long.cu
  return 0;
}

int launch_attention_f16_pack(const float* int_w, const float* int_b, const float* int_weight, const float* int_scaling,
                              int dh, int E, void* pack, cudaStream_t st) {
  EDGL_REQUIRE(attention_f16_pack_bytes(dh, E) != 0, "attention pack: dh=%d E=%d not supported", dh, E);
  if (dh == 16)
    mlp_pack_kernel<16><<<1, 256, 0, st>>>(int_w, int_b, int_weight, int_scaling, reinterpret_cast<unsigned char*>(pack));
  else
    mlp_pack_kernel<32><<<1, 256, 0, st>>>(int_w, int_b, int_weight, int_scaling, reinterpret_cast<unsigned char*>(pack));
  EDGL_LAUNCH_CHECK();
  return 0;
}

// returns 0 = launched, 1 = shape not covered, <0 = error
int launch_attention_f16(const AttnArgs& a, cudaStream_t st) {
  const int dh = a.d / a.h;
  // long sequences / head dim 32: the key-streaming two-pass kernel (attn_f16_long.cu); EDGL_ATTN_LONG=1 forces it
  static const bool force_long = getenv("EDGL_ATTN_LONG") != nullptr;
  if (force_long || dh == 32 || (dh == 16 && a.L > 208)) return launch_attention_f16_long(a, st);
  if (dh != 16 || a.E != 16 || !a.mlp_pack || a.L > 208) return 1;
  if (reinterpret_cast<uintptr_t>(a.marks) & 15) return 1;  // mark rows are read as 16-byte words
  static const int hpc_env = [] {
    const char* e = getenv("EDGL_ATTN_HPC");
    return e ? atoi(e) : 1;
  }();
  static const int occ_env = [] {  // registers per thread of the L <= 104 instantiation (tuning knob)
    const char* e = getenv("EDGL_ATTN_REGS");
    return e ? atoi(e) : 72;
  }();
  int hpc = hpc_env;
  if (hpc < 1 || a.h % hpc != 0) hpc = 1;
  if (a.L <= 32) return launch_f16_t<16, 4, 128>(a, hpc, st);
  if (a.L <= 104) {
    // 72 registers -> 4 CTAs of 7 warps per SM: measured best at C2 (1.34 ms; 80 -> 1.42, 128 -> 1.44)
    if (const int abl = ablation_attn_mask()) {  // precision ablation (tools/ablation.py): the instantiated masks only
      switch (abl) {
        case 3: return launch_f16_t<16, 13, 72, 3>(a, hpc, st);
        case 4: return launch_f16_t<16, 13, 72, 4>(a, hpc, st);
        case 8: return launch_f16_t<16, 13, 72, 8>(a, hpc, st);
        case 12: return launch_f16_t<16, 13, 72, 12>(a, hpc, st);
        case 48: return launch_f16_t<16, 13, 72, 48>(a, hpc, st);
        case 64: return launch_f16_t<16, 13, 72, 64>(a, hpc, st);
        case 127: return launch_f16_t<16, 13, 72, 127>(a, hpc, st);
        default: return set_error(-1, "EDGL_ABL_ATTN=%d is not instantiated (3, 4, 8, 12, 48, 64, 127)", abl);
      }
    }
    if (occ_env == 80) return launch_f16_t<16, 13, 80>(a, hpc, st);
    if (occ_env == 128) return launch_f16_t<16, 13, 128>(a, hpc, st);
    return launch_f16_t<16, 13, 72>(a, hpc, st);
  }
  if (a.L <= 128) return launch_f16_t<16, 16, 128>(a, hpc, st);
  return launch_f16_t<16, 26, 255>(a, hpc, st);
}

}  // namespace edgl
