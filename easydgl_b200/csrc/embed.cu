// embed.cu - input assembly kernels.
// Replaces EasyDGL.__call__ lines EasyDGL.py:70-95 (item gather x sqrt(d) + sinusoid time code,
// position code, mark code, spans, key mask) and CTSMA.py:47-60, plus the stand-alone layers
// C.Embedding (coding.py:45-64) and C.TimeSinusoidCoding (coding.py:125-149).
//
// HBM-bound byte work: one warp per (b,l) row, each lane owns (sin,cos) column pairs so the item
// row is read with coalesced 8-byte loads and every output row is written once, coalesced.
#include "common.cuh"

namespace edgl {

__device__ __forceinline__ float2 ld2(const float* p) { return *reinterpret_cast<const float2*>(p); }
__device__ __forceinline__ void st2(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }

// One warp per (b,l) row; blockDim = 128 (4 warps = 4 consecutive positions of one sequence), grid = (B, ceil(L/4)):
// the kernel is issue-bound (accurate sincos of 64 arguments per row), so (b,l) come from the block index instead of
// a 64-bit division and the mark histogram is done with byte-compare SIMD on the 16-byte mark row.
__global__ void __launch_bounds__(128) embed_kernel(EmbedArgs a, float sqrt_d) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x, l = blockIdx.y * 4 + (threadIdx.x >> 5);
  if (l >= a.L) return;
  const long long row = (long long)b * a.L + l;
  const int d = a.d, E = a.E, half = d >> 1;
  const long long id = a.ids[row];
  const bool id_ok = id > 0 && id < a.num_rows;  // id 0 = zero-padded row (coding.py:56-57)
  const float* ts_row = a.ts + (long long)b * a.ts_len;
  // seqs_t / time_scale: an fp32 divide (EasyDGL.py:71, CTSMA.py:47)
  float span, tsv = 0.f;
  if (a.model == 0) {
    tsv = __fdiv_rn(ts_row[l], a.time_scale);
    const int l1 = l == 0 ? 1 : l;  // spans[0] = spans[1] (EasyDGL.py:74)
    const float t1 = __fdiv_rn(ts_row[l1], a.time_scale), t0 = __fdiv_rn(ts_row[l1 - 1], a.time_scale);
    span = fminf(fmaxf(__fsub_rn(t1, t0), 0.f), 100.f);  // clip_by_value (EasyDGL.py:15-16,73)
  } else {
    const float t1 = __fdiv_rn(ts_row[l + 1], a.time_scale), t0 = __fdiv_rn(ts_row[l], a.time_scale);
    span = __fsub_rn(t1, t0);  // CTSMA.py:49 (unclipped, forward-looking)
  }
  long long mid = (a.model == 0 && id == a.mask_id) ? 0 : id;  // EasyDGL.py:76
  const bool mid_ok = mid >= 0 && mid < a.mark_rows;
  const uint8_t* mrow = a.mark_table8 + (mid_ok ? mid : 0) * E;
  if (lane == 0) {
    if (a.spans) a.spans[row] = span;
    if (a.kmask) a.kmask[row] = id != 0 ? 1 : 0;  // EasyDGL.py:94 / CTSMA.py:59
  }
  if (a.marks)
    for (int e = lane; e < E; e += 32) a.marks[row * E + e] = mid_ok ? mrow[e] : (uint8_t)0;

  const float* irow = a.item_table + (id_ok ? id : 0) * (long long)d;
  float xmax = 0.f;  // max |Xa| of this row (the histogram counts are at most E)
  for (int j = lane; j < half; j += 32) {
    float2 it = id_ok ? ld2(irow + 2 * j) : make_float2(0.f, 0.f);
    float x0 = __fmul_rn(it.x, sqrt_d), x1 = __fmul_rn(it.y, sqrt_d);  // coding.py:61-63
    if (a.model == 0) {
      float s = 0.f, c = 1.f;  // padded slots have ts = 0 -> code [0, 1, 0, 1, ...] (Q10); warp-uniform branch
      if (tsv != 0.f) sincosf(__fdiv_rn(tsv, a.tscale[j]), &s, &c);  // coding.py:142-145 (divide, accurate sin/cos)
      x0 = __fadd_rn(x0, s);                         // EasyDGL.py:83
      x1 = __fadd_rn(x1, c);
    }
    if (a.X0) st2(a.X0 + row * a.ldx0 + 2 * j, x0, x1);
    if (a.Xa) st2(a.Xa + row * a.ldxa + 2 * j, x0, x1);
    xmax = fmaxf(xmax, fmaxf(fabsf(x0), fabsf(x1)));
    if (a.X0) {
      float2 p = ld2(a.pos_table + (long long)l * d + 2 * j);  // coding.py:76-79
      st2(a.X0 + row * a.ldx0 + d + 2 * j, p.x, p.y);
      if (a.model == 0) {
        // mark code: sum_e mark_embs[marks[b,l,e]] with VALUES used as indices (EasyDGL.py:87-88, Q1)
        float m0 = 0.f, m1 = 0.f;
        for (int e = 0; e < E; ++e) {
          const int v = mid_ok ? mrow[e] : 0;
          if (v > 0 && v < E) {  // row 0 of mark_embs is zero-padded
            float2 me = ld2(a.mark_embs + (long long)v * d + 2 * j);
            m0 += me.x;
            m1 += me.y;
          }
        }
        st2(a.X0 + row * a.ldx0 + 2 * d + 2 * j, m0, m1);
      }
    }
  }
  if (a.Xa && a.xa_amax) amax_publish(a.xa_amax, fmaxf(xmax, a.model == 0 ? (float)E : 0.f), lane);
  if (a.Xa && a.model == 0) {
    // histogram of mark values: cnt[v] = #{e : marks[e] == v}; the mark code is cnt @ mark_embs_zp,
    // so the block-0 QKVT dense sees it through a [E,4d] folded kernel (api.cu commit()).
    if (E == 16 && (reinterpret_cast<uintptr_t>(a.mark_table8) & 15) == 0) {
      // the whole mark row is one 16-byte word (same address in every lane: a broadcast load); lane v counts the
      // bytes equal to v with a per-byte compare (0xff per hit -> popcount / 8)
      const uint4 m = mid_ok ? *reinterpret_cast<const uint4*>(mrow) : make_uint4(0u, 0u, 0u, 0u);
      if (lane < 16) {
        const unsigned int vv = (unsigned int)lane * 0x01010101u;
        int c = __popc(__vcmpeq4(m.x, vv)) + __popc(__vcmpeq4(m.y, vv)) + __popc(__vcmpeq4(m.z, vv)) +
                __popc(__vcmpeq4(m.w, vv));
        c >>= 3;
        if (!mid_ok) c = 0;  // TF-GPU semantics for an out-of-range id: zero rows everywhere
        a.Xa[row * a.ldxa + d + lane] = (float)c;
      }
    } else {
      for (int v = lane; v < E; v += 32) {
        int c = 0;
        if (mid_ok)
          for (int e = 0; e < E; ++e) c += (mrow[e] == v);
        a.Xa[row * a.ldxa + d + v] = (float)c;
      }
    }
  }
}

int launch_embed(const EmbedArgs& a, cudaStream_t st) {
  EDGL_REQUIRE(a.d % 2 == 0, "num_units must be even (TimeSinusoidCoding, coding.py:134)");
  EDGL_REQUIRE(a.L >= 2 || a.model == 1, "EasyDGL needs seq_len >= 2 (spans, EasyDGL.py:73-74)");
  const long long rows = (long long)a.B * a.L;
  if (rows == 0) return 0;
  EDGL_REQUIRE(cdiv(a.L, 4) <= 65535, "embed: sequence length %d exceeds the grid limit", a.L);
  embed_kernel<<<dim3((unsigned)a.B, (unsigned)cdiv(a.L, 4)), 128, 0, st>>>(a, (float)sqrt((double)a.d));
  EDGL_LAUNCH_CHECK();
  return 0;
}

__global__ void time_code_kernel(const float* __restrict__ ts, const float* __restrict__ tscale,
                                 long long rows, int d, float* __restrict__ out) {
  const int half = d >> 1;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * half) return;
  const long long row = i / half;
  const int j = (int)(i % half);
  float s, c;
  sincosf(__fdiv_rn(ts[row], tscale[j]), &s, &c);
  st2(out + row * d + 2 * j, s, c);
}

int launch_time_code(const float* ts, const float* tscale, long long rows, int d, float* out, cudaStream_t st) {
  EDGL_REQUIRE(d % 2 == 0 && d > 0, "num_units must be even");
  if (rows == 0) return 0;
  const long long n = rows * (d / 2);
  time_code_kernel<<<cdiv(n, 256), 256, 0, st>>>(ts, tscale, rows, d, out);
  EDGL_LAUNCH_CHECK();
  return 0;
}

// C.TimeFunctionCoding.code (coding.py:112-122): the Bochner / Mercer harmonic time kernel of TGAT,
// out[..., j] = cos(x * basis_freq[j] + phase[j])  (fp32 multiply, then bias_add, then an accurate cos:
// arguments reach 1e5 rad for day-scaled timestamps, so the fast-math cosine is not an option)
__global__ void time_function_kernel(const float* __restrict__ x, const float* __restrict__ freq,
                                     const float* __restrict__ phase, long long n, int d, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * d) return;
  const long long r = i / d;
  const int j = (int)(i % d);
  out[i] = cosf(__fadd_rn(__fmul_rn(x[r], freq[j]), phase[j]));
}

int launch_time_function_code(const float* x, const float* freq, const float* phase, long long n, int d, float* out,
                              cudaStream_t st) {
  EDGL_REQUIRE(d > 0, "num_units must be positive");
  if (n == 0) return 0;
  time_function_kernel<<<cdiv(n * d, 256), 256, 0, st>>>(x, freq, phase, n, d, out);
  EDGL_LAUNCH_CHECK();
  return 0;
}

__global__ void lookup_kernel(const float* __restrict__ table, int vocab, int d, int zero_pad, float scale,
                              const int64_t* __restrict__ ids, long long n, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * d) return;
  const long long r = i / d;
  const int c = (int)(i % d);
  const long long id = ids[r];
  float v = 0.f;
  if (id >= 0 && id < vocab && !(zero_pad && id == 0)) v = table[id * d + c];
  out[i] = __fmul_rn(v, scale);
}

int launch_lookup(const float* table, int vocab, int d, int zero_pad, float scale, const int64_t* ids,
                  long long n, float* out, cudaStream_t st) {
  if (n == 0) return 0;
  lookup_kernel<<<cdiv(n * d, 256), 256, 0, st>>>(table, vocab, d, zero_pad, scale, ids, n, out);
  EDGL_LAUNCH_CHECK();
  return 0;
}

__global__ void mark_u8_kernel(const int64_t* __restrict__ src, long long n, uint8_t* __restrict__ dst,
                               int* err_flag, int E) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long v = src[i];
  if (v < 0 || v >= E) atomicExch(err_flag, 1);
  dst[i] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

int launch_mark_table_to_u8(const int64_t* src, long long n, uint8_t* dst, int* err_flag, int E,
                            cudaStream_t st) {
  if (n == 0) return 0;
  mark_u8_kernel<<<cdiv(n, 256), 256, 0, st>>>(src, n, dst, err_flag, E);
  EDGL_LAUNCH_CHECK();
  return 0;
}

}  // namespace edgl
