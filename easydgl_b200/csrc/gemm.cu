// gemm.cu - fp32 dense layer with fused epilogue.
// Replaces every tf.layers.dense / Conv1D(1) / tied-logits tf.matmul on the hot path
// (temporal.py:340-343,409; EasyDGL.py:113,120,125,138,149-150; Base.py:77-87; CTSMA.py:89-90).
//
// v1 arithmetic is exact fp32 FMA (CUDA cores): the ranking must match an fp32 reference, and plain
// TF32 rounding (2^-11 per operand) flips top-K sets (hard part 1).
// 128x128x16 tiles, 256 threads, 8x8 register micro-tiles, register-staged double buffering.
// Epilogue: + bias[N] + periodic bias[(m % period), N] -> activation -> + residual.
#include <stdlib.h>

#include "common.cuh"

namespace edgl {

constexpr int BM = 128, BN = 128, BK = 16, PAD = 4;

__device__ __forceinline__ float gelu_erf(float x) {
  // EasyDGL.gelu (EasyDGL.py:31-32): x * 0.5 * (1 + erf(x / sqrt(2)))
  const float cdf = 0.5f * (1.0f + erff(__fdiv_rn(x, 1.41421356237309504880f)));
  return x * cdf;
}

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == ACT_GELU) return gelu_erf(v);
  if (act == ACT_RELU) return fmaxf(v, 0.f);
  return v;
}

// load a [rows x 4] strip: 4 consecutive k of one row, guarded
__device__ __forceinline__ float4 load_k4(const float* __restrict__ base, long long row_off, int k, int K,
                                          bool row_ok, bool vec_ok) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (!row_ok) return v;
  const float* p = base + row_off + k;
  if (vec_ok) {
    if (k < K) v = *reinterpret_cast<const float4*>(p);
  } else {
    if (k + 0 < K) v.x = p[0];
    if (k + 1 < K) v.y = p[1];
    if (k + 2 < K) v.z = p[2];
    if (k + 3 < K) v.w = p[3];
  }
  return v;
}

template <bool W_NK>
__global__ void __launch_bounds__(256, 2) gemm_f32_kernel(GemmArgs a, int ntn, bool a_vec, bool w_vec) {
  __shared__ __align__(16) float As[2][BK][BM + PAD];
  __shared__ __align__(16) float Bs[2][BK][BN + PAD];
  const int t = threadIdx.x;
  const int tile = blockIdx.x;
  const int n0 = (tile % ntn) * BN;
  const long long m0 = (long long)(tile / ntn) * BM;
  const int M = a.M, N = a.N, K = a.K;

  // A loader mapping: rows ar, ar+64; k offset ak
  const int ar = t >> 2, ak = (t & 3) * 4;
  // W loader mapping ([K,N]): k rows wk, wk+8; n offset wn. ([N,K]): like A.
  const int wk = t >> 5, wn = (t & 31) * 4;

  float4 ra[2], rw[2];
  auto gload = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const long long r = m0 + ar + 64 * i;
      ra[i] = load_k4(a.A, r * a.lda, k0 + ak, K, r < M, a_vec);
    }
    if (W_NK) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const long long n = (long long)n0 + ar + 64 * i;
        const bool ok = n < N && !(a.zero_wrow0 && n == 0);
        rw[i] = load_k4(a.W, n * a.ldw, k0 + ak, K, ok, w_vec);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int k = k0 + wk + 8 * i;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k < K) {
          const float* p = a.W + (long long)k * a.ldw + n0 + wn;
          if (w_vec) {
            if (n0 + wn < N) v = *reinterpret_cast<const float4*>(p);
          } else {
            if (n0 + wn + 0 < N) v.x = p[0];
            if (n0 + wn + 1 < N) v.y = p[1];
            if (n0 + wn + 2 < N) v.z = p[2];
            if (n0 + wn + 3 < N) v.w = p[3];
          }
        }
        rw[i] = v;
      }
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      As[buf][ak + 0][ar + 64 * i] = ra[i].x;
      As[buf][ak + 1][ar + 64 * i] = ra[i].y;
      As[buf][ak + 2][ar + 64 * i] = ra[i].z;
      As[buf][ak + 3][ar + 64 * i] = ra[i].w;
    }
    if (W_NK) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        Bs[buf][ak + 0][ar + 64 * i] = rw[i].x;
        Bs[buf][ak + 1][ar + 64 * i] = rw[i].y;
        Bs[buf][ak + 2][ar + 64 * i] = rw[i].z;
        Bs[buf][ak + 3][ar + 64 * i] = rw[i].w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 2; ++i) *reinterpret_cast<float4*>(&Bs[buf][wk + 8 * i][wn]) = rw[i];
    }
  };

  const int ty = t >> 4, tx = t & 15;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  const int nk = (K + BK - 1) / BK;
  gload(0);
  sstore(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) gload((kt + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      sstore(buf ^ 1);
      __syncthreads();
    }
  }

  // epilogue
  const bool c_vec = (a.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.C) & 15) == 0) && (N % 4 == 0) &&
                     (a.R == nullptr || ((a.ldr % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.R) & 15) == 0)));
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long long row = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (row >= M) continue;
    const float* pb = a.pbias ? a.pbias + (row % a.pperiod) * (long long)N : nullptr;
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) {
      const int col = n0 + jj * 64 + tx * 4;
      if (col >= N) continue;
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float x = acc[i][jj * 4 + j];
        const int c = col + j;
        if (c < N) {
          if (a.bias) x += a.bias[c];
          if (pb) x += pb[c];
          x = apply_act(x, a.act);
          if (a.R) x += a.R[row * a.ldr + c];
        }
        v[j] = x;
      }
      float* pc = a.C + row * a.ldc + col;
      if (c_vec) {
        *reinterpret_cast<float4*>(pc) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (col + j < N) pc[j] = v[j];
      }
    }
  }
}

// out[c][r] = in[r][c]  (weights are transposed once at commit so the tensor-core kernel sees K-major W)
__global__ void transpose_kernel(const float* __restrict__ in, int rows, int cols, float* __restrict__ out) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? in[(long long)r * cols + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < rows && c < cols) out[(long long)c * rows + r] = tile[threadIdx.x][i];
  }
}

int launch_transpose(const float* in, int rows, int cols, float* out, cudaStream_t st) {
  if (rows == 0 || cols == 0) return 0;
  transpose_kernel<<<dim3(cdiv(cols, 32), cdiv(rows, 32)), dim3(32, 8), 0, st>>>(in, rows, cols, out);
  EDGL_LAUNCH_CHECK();
  return 0;
}

__global__ void tf32_lo_kernel(const float* __restrict__ in, long long n, float* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float x = in[i];
    out[i] = x - __uint_as_float(__float_as_uint(x) & 0xffffe000u);
  }
}

int launch_tf32_lo(const float* in, long long n, float* out, cudaStream_t st) {
  if (n == 0) return 0;
  long long blocks = (n + 255) / 256;
  if (blocks > 4096) blocks = 4096;
  tf32_lo_kernel<<<(unsigned)blocks, 256, 0, st>>>(in, n, out);
  EDGL_LAUNCH_CHECK();
  return 0;
}

int launch_gemm(const GemmArgs& a, cudaStream_t st) {
  EDGL_REQUIRE(a.M >= 0 && a.N > 0 && a.K > 0, "gemm: bad shape M=%d N=%d K=%d", a.M, a.N, a.K);
  if (a.M == 0) return 0;
  EDGL_REQUIRE(!a.pbias || a.pperiod > 0, "gemm: periodic bias needs a period");
  // Blackwell tensor-core path (gemm_tc.cu) whenever W is K-major; EDGL_GEMM=simt forces the CUDA-core kernel
  static const bool force_simt = [] {
    const char* e = getenv("EDGL_GEMM");
    return e && e[0] == 's';
  }();
  if (a.flt.cand || a.run_if) {  // candidate filter / predicated launch: 3xTF32 tensor-core kernel only
    EDGL_REQUIRE(gemm_tc_supported(a), "gemm: the top-K filter epilogue needs the tensor-core path");
    return launch_gemm_tc(a, st);
  }
  if (a.ln.any()) {  // fused LayerNorm pieces exist in the 3xTF32 tensor-core kernel only (callers check EDGL_GEMM)
    EDGL_REQUIRE(gemm_tc_supported(a), "gemm: a fused-LayerNorm dense layer needs the tensor-core path");
    return launch_gemm_tc(a, st);
  }
  if (!force_simt && gemm_f16_supported(a)) return launch_gemm_f16(a, st);  // scaled 3xFP16 on kind::f16
  if (!force_simt && gemm_tc_supported(a)) return launch_gemm_tc(a, st);
  const int ntn = cdiv(a.N, BN);
  const long long ntm = cdiv(a.M, BM);
  EDGL_REQUIRE(ntm * ntn < (1ll << 31), "gemm: grid too large");
  const bool a_vec = (a.lda % 4 == 0) && (a.K % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.A) & 15) == 0);
  bool w_vec;
  if (a.w_is_nk)
    w_vec = (a.ldw % 4 == 0) && (a.K % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.W) & 15) == 0);
  else
    w_vec = (a.ldw % 4 == 0) && (a.N % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.W) & 15) == 0);
  const unsigned grid = (unsigned)(ntm * ntn);
  if (a.w_is_nk)
    gemm_f32_kernel<true><<<grid, 256, 0, st>>>(a, ntn, a_vec, w_vec);
  else
    gemm_f32_kernel<false><<<grid, 256, 0, st>>>(a, ntn, a_vec, w_vec);
  EDGL_LAUNCH_CHECK();
  return 0;
}

}  // namespace edgl
