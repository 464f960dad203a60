// gemm_f16.cu - the dense layers on tcgen05.mma kind::f16 with the SCALED 3xFP16 split of attn_f16.cu.
//
// Same role, pipeline and epilogue as gemm_tc.cu (tf.layers.dense / tied-logits matmuls: temporal.py:409;
// EasyDGL.py:113,120,125,138,149), but the three tensor-core products per k-step run on fp16 operands:
//     C = (A_lo*W_hi + A_hi*W_lo + A_hi*W_hi) / (sa*sw),   x*s = hi + lo,  hi = top 11 significant bits (exact in fp16)
// The dense GEMMs are bound by the shared-memory port: fp16 operands halve the bytes every MMA reads
// and the bytes TMA writes for W, and each MMA covers k = 16 instead of 8.
//   * W: split once per edgl_commit into fp16 hi / lo copies with a per-tensor power-of-two scale sw
//     (launch_w_split_f16); both arrive by TMA (64-byte swizzle).
//   * A: activations arrive as fp32 by TMA (128-byte swizzle) exactly as in gemm_tc.cu; the four splitter warps
//     convert each k-block to fp16 hi / lo tiles (64-byte swizzle, written in the layout the MMA descriptor expects)
//     with the per-tensor scale sa = 2^k that puts max|A| into [2^14, 2^15).  max|A| is published by the kernel
//     that PRODUCED A (amax_publish in embed / attention / layernorm / this epilogue), so no extra pass reads A.
//     Elements up to 2^18 below the tensor's maximum keep full relative precision (fp16 subnormals give an absolute
//     floor of 2^-25 after scaling), i.e. the error relative to max|A|*|W| is that of the 3xTF32 kernel.
//   * the scales are exact powers of two and are divided out of the fp32 accumulator in the epilogue.
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include "tc_common.cuh"

namespace edgl {

namespace tcf {

using namespace tcc;

// BK (tc_common.cuh) = 32 elements per k-block: one 128-byte swizzle row of fp32 A, one 64-byte row of fp16
constexpr int NTHREADS = TC_THREADS;

// K-major, 64B-swizzled shared-memory operand descriptor (rows of 64 B = 32 halves, 8-row atoms 512 B apart)
__device__ __forceinline__ uint64_t umma_desc64(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3ffff) >> 4);       // start address >> 4, bits [0,14)
  d |= (uint64_t)1 << 16;                        // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(512 >> 4) << 32;               // stride byte offset: 8 rows * 64 B
  d |= (uint64_t)1 << 46;                        // descriptor version (Blackwell)
  d |= (uint64_t)4 << 61;                        // SWIZZLE_64B
  return d;
}
// instruction descriptor: D=f32, A=B=f16 (format 0), both K-major, M=128, N=bn
__host__ __device__ constexpr uint32_t umma_idesc(int bn) {
  return (1u << 4) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// m >= 0: s = 2^k with m*s in [2^14, 2^15), is = 1/s (exact; exponent clamped so neither is denormal)
__device__ __forceinline__ void pow2_scale(uint32_t mbits, float& s, float& is) {
  int e = (int)((mbits >> 23) & 0xffu);
  e = e < 15 ? 15 : (e > 239 ? 239 : e);
  s = __uint_as_float((uint32_t)(268 - e) << 23);
  is = __uint_as_float((uint32_t)(e - 14) << 23);
}

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

struct Params {
  int gelu_fit;  // GELU epilogue: 1 = gelu_fit (default), 0 = exact erff (EDGL_GELU=erf)
  float* C; int ldc;
  int M, N, K;
  const float* bias;
  const float* pbias; int pperiod;
  const float* R; int ldr;
  int act;
  int col0_bias_only;
  int ntn, num_tiles, kblocks;
  int epi_direct;
  const unsigned int* a_amax;  // bits of max|A| (published by A's producer)
  const float* w_inv;          // 1 / sw
  unsigned int* c_amax;        // optional: publish max|C| for the next layer
  int abl;                     // precision ablation bits (common.cuh), 0 on the product path
};

template <int BN>
struct Smem {
  static constexpr int A32_BYTES = BM * BK * 4;   // 16 KB fp32 A k-block as landed by TMA (SW128)
  static constexpr int AH_BYTES = BM * BK * 2;    // 8 KB fp16 hi (and lo) tile (SW64)
  static constexpr int B_BYTES = BN * BK * 2;     // fp16 W hi (and lo) tile (SW64)
  static constexpr int STAGE = A32_BYTES + 2 * AH_BYTES + 2 * B_BYTES;
  static constexpr int STAGES = (BN == 256) ? 3 : 4;
  static constexpr int BYTES = STAGES * STAGE + 1024 /*align*/ + 256 /*barriers*/ + EPI_WARPS * 32 * CP * 4 /*epilogue*/;
};

template <int BN, int ACT>
__global__ void __launch_bounds__(NTHREADS, 1)
gemm_f16_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapBh,
                const __grid_constant__ CUtensorMap mapBl, Params p) {
  using SM = Smem<BN>;
  constexpr int S = SM::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + S * SM::STAGE);
  uint64_t* full = bars;            // [S]
  uint64_t* split = bars + S;       // [S]
  uint64_t* empty = bars + 2 * S;   // [S]
  uint64_t* tfull = bars + 3 * S;   // [2]
  uint64_t* tempty = tfull + 2;     // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  auto stA32 = [&](int s) { return base + s * SM::STAGE; };
  auto stAh = [&](int s) { return base + s * SM::STAGE + SM::A32_BYTES; };
  auto stAl = [&](int s) { return base + s * SM::STAGE + SM::A32_BYTES + SM::AH_BYTES; };
  auto stBh = [&](int s) { return base + s * SM::STAGE + SM::A32_BYTES + 2 * SM::AH_BYTES; };
  auto stBl = [&](int s) { return base + s * SM::STAGE + SM::A32_BYTES + 2 * SM::AH_BYTES + SM::B_BYTES; };

  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&split[i], 128);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], EPI_WARPS * 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(2 * BN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // activation scale from the producer's running maximum (the producing kernel has completed: stream order)
  float sa, isa;
  pow2_scale(*p.a_amax, sa, isa);

  const int KB = p.kblocks;
  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapA)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapBh)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapBl)) : "memory");
      uint32_t it = 0;
      for (int kb = 0; kb < KB; ++kb) tma_prefetch_2d(&mapA, kb * BK, (blockIdx.x / p.ntn) * BM);
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int m0 = (tile / p.ntn) * BM, n0 = (tile % p.ntn) * BN;
        const int ntile = tile + gridDim.x;
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const int s = it % S;
          if (ntile < p.num_tiles) tma_prefetch_2d(&mapA, kb * BK, (ntile / p.ntn) * BM);
          mbar_wait(&empty[s], ((it / S) & 1) ^ 1);
          mbar_expect_tx(&full[s], SM::A32_BYTES + 2 * SM::B_BYTES);
          tma_load_2d(&mapA, &full[s], stA32(s), kb * BK, m0);
          tma_load_2d(&mapBh, &full[s], stBh(s), kb * BK, n0);
          tma_load_2d(&mapBl, &full[s], stBl(s), kb * BK, n0);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc(BN);
      uint32_t it = 0, tcount = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tcount) {
        const int acc = tcount & 1;
        mbar_wait(&tempty[acc], ((tcount >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const int s = it % S;
          mbar_wait(&split[s], (it / S) & 1);
          tc_fence_after();
          const uint32_t a_hi = smem_u32(stAh(s)), a_lo = smem_u32(stAl(s));
          const uint32_t b_hi = smem_u32(stBh(s)), b_lo = smem_u32(stBl(s));
#pragma unroll
          for (int k = 0; k < 2; ++k) {          // two k16 steps per 32-element k-block
            const uint32_t off = k * 32;         // bytes along K inside the 64 B swizzle row
            const uint64_t dah = umma_desc64(a_hi + off), dal = umma_desc64(a_lo + off);
            const uint64_t dbh = umma_desc64(b_hi + off), dbl = umma_desc64(b_lo + off);
            uint32_t accum = (kb | k) != 0;  // p.abl: precision ablation (common.cuh), 0 on the product path
            if (!(p.abl & 1)) { umma_f16(d_tmem, dal, dbh, idesc, accum); accum = 1; }
            if (!(p.abl & 2)) { umma_f16(d_tmem, dah, dbl, idesc, accum); accum = 1; }
            umma_f16(d_tmem, dah, dbh, idesc, accum);
          }
          umma_commit(&empty[s]);          // frees the stage when these MMAs have read it
        }
        umma_commit(&tfull[acc]);          // accumulator complete
      }
    }
  } else if (warp >= SPLIT_WARP0) {
    // ------------------------------------------------------------------ splitters: fp32 (SW128) -> fp16 hi, lo (SW64)
    // thread = tile row.  Source row r: 128 B at (r/8)*1024 + (r%8)*128, 16-byte chunk c stored at c ^ (r%8).
    // Destination row r: 64 B at (r/8)*512 + (r%8)*64, 16-byte chunk c4 (8 halves) stored at c4 ^ ((r/2)%4).
    const int r = threadIdx.x - SPLIT_WARP0 * 32;  // 0..127
    const uint32_t src_row = (r >> 3) * 1024 + (r & 7) * 128, sx = r & 7;
    const uint32_t dst_row = (r >> 3) * 512 + (r & 7) * 64, dx = (r >> 1) & 3;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      for (int kb = 0; kb < KB; ++kb, ++it) {
        const int s = it % S;
        mbar_wait(&full[s], (it / S) & 1);
        const uint8_t* a32 = stA32(s);
        uint8_t* ah = stAh(s);
        uint8_t* al = stAl(s);
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          const float4 v0 = *reinterpret_cast<const float4*>(a32 + src_row + (((2 * c4) ^ sx) << 4));
          const float4 v1 = *reinterpret_cast<const float4*>(a32 + src_row + (((2 * c4 + 1) ^ sx) << 4));
          uint4 hi, lo;
          split2(v0.x * sa, v0.y * sa, hi.x, lo.x);
          split2(v0.z * sa, v0.w * sa, hi.y, lo.y);
          split2(v1.x * sa, v1.y * sa, hi.z, lo.z);
          split2(v1.z * sa, v1.w * sa, hi.w, lo.w);
          *reinterpret_cast<uint4*>(ah + dst_row + ((c4 ^ dx) << 4)) = hi;
          *reinterpret_cast<uint4*>(al + dst_row + ((c4 ^ dx) << 4)) = lo;
        }
        fence_proxy_async();  // generic-proxy writes -> visible to the tensor core (async proxy)
        mbar_arrive(&split[s]);
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue (see gemm_tc.cu)
    const int q = warp & 3;            // TMEM lane quarter this warp may read
    const int half = (warp - 4) >> 2;  // the EPI_PARTS warps of a quarter take every EPI_PARTS-th 16-column sub-chunk
    float* stg = reinterpret_cast<float*>(base + S * SM::STAGE + 256) + (warp - 4) * (32 * CP);
    const bool all_al = (p.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) &&
                        (!p.R || ((p.ldr % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.R) & 15) == 0))) &&
                        (!p.bias || ((reinterpret_cast<uintptr_t>(p.bias) & 15) == 0)) &&
                        (!p.pbias || ((p.N % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.pbias) & 15) == 0)));
    constexpr int LPR = CW / 4, RPI = 32 / LPR, NIT = 32 / RPI;
    const int rsub = lane / LPR;  // this lane's row inside a group of RPI (staged path)
    const float inv = isa * (*p.w_inv);  // accumulator -> A @ W
    float cmax = 0.f;                    // max |C| written by this thread
    uint32_t tcount = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tcount) {
      const int acc = tcount & 1;
      const int m0 = (tile / p.ntn) * BM, n0 = (tile % p.ntn) * BN;
      const int rbase = m0 + q * 32 + rsub;
      int pbo[NIT];
      if (p.pbias) {
        int pr = rbase % p.pperiod;
        const int step = RPI % p.pperiod;
#pragma unroll
        for (int itr = 0; itr < NIT; ++itr) {
          pbo[itr] = pr * p.N;
          pr += step;
          if (pr >= p.pperiod) pr -= p.pperiod;
        }
      }
      int pbd[4] = {0, 0, 0, 0};
      if (p.pbias && p.epi_direct) {
#pragma unroll
        for (int i = 0; i < 4; ++i) pbd[i] = ((m0 + q * 32 + i * 8 + (lane >> 2)) % p.pperiod) * p.N;
      }
      mbar_wait(&tfull[acc], (tcount >> 1) & 1);
      tc_fence_after();
      epilogue_tile<BN, ACT, true>(p, tmem_base + acc * BN, m0, n0, q, half, lane, stg, all_al, pbo, pbd, inv, cmax);
      tc_fence_before();
      mbar_arrive(&tempty[acc]);
    }
    __syncwarp();
    if (p.c_amax) amax_publish(p.c_amax, cmax, lane);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * BN));
  }
}

// ---- weights: per-tensor maximum, then fp16 hi / lo copies and 1/scale
__global__ void w_absmax_kernel(const float* __restrict__ w, long long n, unsigned int* __restrict__ slot) {
  float m = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(w[i]));
  amax_publish(slot, m, threadIdx.x & 31);
}
__global__ void w_split_kernel(const float* __restrict__ w, long long n, const unsigned int* __restrict__ slot,
                               __half* __restrict__ hi, __half* __restrict__ lo, float* __restrict__ inv) {
  float s, is;
  pow2_scale(*slot, s, is);
  if (blockIdx.x == 0 && threadIdx.x == 0) *inv = is;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float x = w[i] * s;
    const float h = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
    hi[i] = __float2half_rn(h);
    lo[i] = __float2half_rn(x - h);
  }
}

}  // namespace tcf

// hi / lo / inv / scratch live in one caller-provided buffer of w16_bytes(n): [hi n halves][lo n halves][inv f32][max u32]
size_t w16_bytes(long long n) { return (size_t)n * 4 + 32; }

int launch_w_split_f16(const float* w, long long n, void* buf, cudaStream_t st) {
  if (n <= 0) return 0;
  __half* hi = reinterpret_cast<__half*>(buf);
  __half* lo = hi + n;
  float* inv = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(buf) + (size_t)n * 4);
  unsigned int* slot = reinterpret_cast<unsigned int*>(inv) + 4;
  EDGL_CUDA(cudaMemsetAsync(slot, 0, sizeof(unsigned int), st));
  long long blocks = (n + 255) / 256;
  if (blocks > 2048) blocks = 2048;
  tcf::w_absmax_kernel<<<(unsigned)blocks, 256, 0, st>>>(w, n, slot);
  EDGL_LAUNCH_CHECK();
  tcf::w_split_kernel<<<(unsigned)blocks, 256, 0, st>>>(w, n, slot, hi, lo, inv);
  EDGL_LAUNCH_CHECK();
  return 0;
}

// running max |x| of a whole tensor into `slot` (which the caller has zeroed): the stand-alone entry point's
// substitute for a producer kernel's amax_publish
int launch_absmax(const float* x, long long n, unsigned int* slot, cudaStream_t st) {
  if (n <= 0) return 0;
  long long blocks = (n + 255) / 256;
  if (blocks > 2048) blocks = 2048;
  tcf::w_absmax_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, n, slot);
  EDGL_LAUNCH_CHECK();
  return 0;
}

bool gemm_f16_supported(const GemmArgs& a) {
  if (!a.w_is_nk || !a.W16 || !a.a_amax) return false;
  if (a.M < 1 || a.K < 8 || (a.K % 8)) return false;   // fp16 rows: 16-byte pitch
  if ((a.lda % 4) || a.ldw != a.K) return false;
  if ((reinterpret_cast<uintptr_t>(a.A) & 15) || (reinterpret_cast<uintptr_t>(a.W16) & 15)) return false;
  return true;
}

int launch_gemm_f16(const GemmArgs& a, cudaStream_t st) {
  using namespace tcf;
  static int num_sms = [] {
    int dev = 0, n = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n;
  }();
  if (a.M == 0) return 0;
  const int bn = a.N > 128 ? 256 : 128;
  const long long nw = (long long)a.N * a.K;
  const __half* wh = reinterpret_cast<const __half*>(a.W16);
  const __half* wl = wh + nw;
  const float* winv = reinterpret_cast<const float*>(reinterpret_cast<const unsigned char*>(a.W16) + (size_t)nw * 4);
  CUtensorMap mapA, mapBh, mapBl;
  EDGL_TRY(make_map(&mapA, a.A, false, a.M, a.K, a.lda, BM));
  EDGL_TRY(make_map(&mapBh, wh, true, a.N, a.K, a.K, bn));
  EDGL_TRY(make_map(&mapBl, wl, true, a.N, a.K, a.K, bn));
  Params p;
  p.C = a.C; p.ldc = a.ldc; p.M = a.M; p.N = a.N; p.K = a.K; p.bias = a.bias; p.pbias = a.pbias;
  p.pperiod = a.pperiod > 0 ? a.pperiod : 1; p.R = a.R; p.ldr = a.ldr; p.act = a.act;
  p.col0_bias_only = a.zero_wrow0 ? 1 : 0;
  static const bool gelu_exact = [] { const char* e = getenv("EDGL_GELU"); return e && e[0] == 'e'; }();
  p.gelu_fit = gelu_exact ? 0 : 1;
  p.a_amax = a.a_amax; p.w_inv = winv; p.c_amax = a.c_amax;
  p.abl = ablation_gemm_bits();
  static const char epi_mode = [] {
    const char* e = getenv("EDGL_TC_EPI");
    return e ? e[0] : 'a';
  }();
  p.epi_direct = epi_mode == 'd' ? 1 : epi_mode == 's' ? 0 : (a.R == nullptr && a.pbias == nullptr) ? 1 : 0;
  p.ntn = cdiv(a.N, bn);
  const long long ntm = cdiv(a.M, BM);
  EDGL_REQUIRE(ntm * p.ntn < (1ll << 31), "gemm_f16: too many tiles");
  p.num_tiles = (int)(ntm * p.ntn);
  p.kblocks = cdiv(a.K, BK);
  const int grid = p.num_tiles < num_sms ? p.num_tiles : num_sms;
#define EDGL_F16_LAUNCH(BNV, ACTV)                                                                             \
  {                                                                                                            \
    auto kern = gemm_f16_kernel<BNV, ACTV>;                                                                    \
    EDGL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<BNV>::BYTES));      \
    kern<<<grid, NTHREADS, Smem<BNV>::BYTES, st>>>(mapA, mapBh, mapBl, p);                                     \
  }
  if (bn == 256) {
    if (a.act == ACT_GELU) EDGL_F16_LAUNCH(256, ACT_GELU)
    else if (a.act == ACT_RELU) EDGL_F16_LAUNCH(256, ACT_RELU)
    else EDGL_F16_LAUNCH(256, ACT_NONE)
  } else {
    if (a.act == ACT_GELU) EDGL_F16_LAUNCH(128, ACT_GELU)
    else if (a.act == ACT_RELU) EDGL_F16_LAUNCH(128, ACT_RELU)
    else EDGL_F16_LAUNCH(128, ACT_NONE)
  }
#undef EDGL_F16_LAUNCH
  EDGL_LAUNCH_CHECK();
  return 0;
}

}  // namespace edgl
