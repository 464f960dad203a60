// gemm_tc.cu - Blackwell-native dense layer: tcgen05.mma (kind::tf32, accumulators in TMEM), operands
// staged by TMA (128B swizzle) through an mbarrier pipeline, warp-specialised persistent CTAs.
//
// Replaces the tf.layers.dense / tied-logits matmuls of the hot path (temporal.py:409, 340-343;
// EasyDGL.py:113,120,125,138,149; Base.py:77-87).  Arithmetic is 3xTF32:
//     C = A_lo*W_hi + A_hi*W_lo + A_hi*W_hi,    x_hi = the 19 bits the tensor core reads, x_lo = x - x_hi
// so the result carries fp32-level accuracy (the ranking parity bar rules out plain TF32).
// The lo halves are produced on the fly in shared memory by four "splitter" warps (an element-wise pass
// over the freshly landed TMA tiles - the swizzled layout is preserved because the op is element-wise),
// so HBM/L2 only ever carry the fp32 operands once.
//
// Roles (640 threads, 1 CTA/SM, persistent over 128 x BN output tiles):
//   warp 0        TMA producer            warp 1       tcgen05.mma issuer (one elected lane)
//   warp 2        TMEM allocator          warps 4-15   epilogue: tcgen05.ld -> bias/act/residual -> global
//   warps 16-19   splitters (A_lo, W_lo)               (three epilogue warps per TMEM lane quarter)
// Pipelines: full[s] (TMA landed) -> split[s] (lo halves written) -> MMA -> empty[s];
//            tmem_full[a] (tile accumulated) -> epilogue -> tmem_empty[a]  (two TMEM accumulators).
#include <cuda.h>
#include <stdlib.h>

#include "tc_common.cuh"

namespace edgl {

namespace tc {

using namespace tcc;

// BK (tc_common.cuh) = 32 floats per k-block = one 128-byte swizzle row
constexpr int UK = 8;        // UMMA K for tf32 (32 bytes)
constexpr int NTHREADS = TC_THREADS;

// K-major, 128B-swizzled shared-memory operand descriptor (rows of 128 B, 8-row groups 1024 B apart)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3ffff) >> 4);       // start address >> 4, bits [0,14)
  d |= (uint64_t)1 << 16;                        // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;              // stride byte offset: 8 rows * 128 B
  d |= (uint64_t)1 << 46;                        // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                        // SWIZZLE_128B
  return d;
}
// instruction descriptor: D=f32, A=B=tf32, both K-major, M=128, N=bn
__host__ __device__ constexpr uint32_t umma_idesc(int bn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct Params {
  int gelu_fit;  // GELU epilogue: 1 = gelu_fit (default), 0 = exact erff (EDGL_GELU=erf)
  float* C; int ldc;
  int M, N, K;
  const float* bias;
  const float* pbias; int pperiod;
  const float* R; int ldr;
  int act;
  int col0_bias_only;  // tied zero-padded table: column 0 is exactly the bias (coding.py:56-57)
  int ntn, num_tiles, kblocks;
  LnEpi ln;        // fused LayerNorm pieces of the epilogue (common.cuh); all null = off
  TopkFilter flt;  // top-K candidate filter instead of the store (LN_FILT instantiation only)
  const int* run_if;  // device flag or null: return at once when *run_if == 0
  int abl;            // precision ablation bits (common.cuh), 0 on the product path
  int epi_direct;  // epilogue stores rows straight from TMEM fragments (no shared-memory transpose)
  int has_blo;  // W_lo = W - tf32(W) is pre-computed in global memory (weights are constant after commit): TMA brings
                // it in like W and the splitter warps only split the activations
  long long* dbg;  // optional [gridDim][8] cycle counters (EDGL_TC_DEBUG), else null
};

template <int BN>
struct Smem {
  static constexpr int A_BYTES = BM * BK * 4;     // 16 KB
  static constexpr int B_BYTES = BN * BK * 4;
  static constexpr int STAGE = 2 * A_BYTES + 2 * B_BYTES;
  static constexpr int STAGES = (BN == 256) ? 2 : 3;
  static constexpr int LN_BYTES = 2 * 256 * 4;  // gamma / beta of a LayerNorm folded into A (K <= 256)
  static constexpr int BYTES = STAGES * STAGE + 1024 /*align*/ + 256 /*barriers*/ + EPI_WARPS * 32 * CP * 4 /*epilogue*/ + LN_BYTES;
};

template <int BN, int ACT, int LNF>
__global__ void __launch_bounds__(NTHREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
               const __grid_constant__ CUtensorMap mapBlo, Params p) {
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

  if (p.run_if && *p.run_if == 0) return;  // predicated fallback launch (block-uniform; nothing is allocated yet)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  auto stA = [&](int s) { return base + s * SM::STAGE; };
  auto stAlo = [&](int s) { return base + s * SM::STAGE + SM::A_BYTES; };
  auto stB = [&](int s) { return base + s * SM::STAGE + 2 * SM::A_BYTES; };
  auto stBlo = [&](int s) { return base + s * SM::STAGE + 2 * SM::A_BYTES + SM::B_BYTES; };

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

  const int KB = p.kblocks;
  long long dbg_acc[7] = {0, 0, 0, 0, 0, 0, 0};
  const long long t_start = clock64();
  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapA)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapB)) : "memory");
      if (p.has_blo) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapBlo)) : "memory");
      uint32_t it = 0;
      // the activations stream from HBM exactly once: pull the A rows of the NEXT tile into L2 while the
      // current tile is computed, so the TMA loads below see L2 latency instead of DRAM latency
      for (int kb = 0; kb < KB; ++kb) tma_prefetch_2d(&mapA, kb * BK, (blockIdx.x / p.ntn) * BM);
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int m0 = (tile / p.ntn) * BM, n0 = (tile % p.ntn) * BN;
        const int ntile = tile + gridDim.x;
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const int s = it % S;
          if (ntile < p.num_tiles) tma_prefetch_2d(&mapA, kb * BK, (ntile / p.ntn) * BM);
          const long long t0 = clock64();
          mbar_wait(&empty[s], ((it / S) & 1) ^ 1);
          dbg_acc[0] += clock64() - t0;
          mbar_expect_tx(&full[s], SM::A_BYTES + (p.has_blo ? 2 * SM::B_BYTES : SM::B_BYTES));
          tma_load_2d(&mapA, &full[s], stA(s), kb * BK, m0);
          tma_load_2d(&mapB, &full[s], stB(s), kb * BK, n0);
          if (p.has_blo) tma_load_2d(&mapBlo, &full[s], stBlo(s), kb * BK, n0);
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
        long long t0 = clock64();
        mbar_wait(&tempty[acc], ((tcount >> 1) & 1) ^ 1);
        dbg_acc[1] += clock64() - t0;
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const int s = it % S;
          t0 = clock64();
          mbar_wait(&split[s], (it / S) & 1);
          dbg_acc[2] += clock64() - t0;
          tc_fence_after();
          const uint32_t a_hi = smem_u32(stA(s)), a_lo = smem_u32(stAlo(s));
          const uint32_t b_hi = smem_u32(stB(s)), b_lo = smem_u32(stBlo(s));
#pragma unroll
          for (int k = 0; k < BK / UK; ++k) {
            const uint32_t off = k * UK * 4;  // bytes along K inside the 128 B swizzle row
            const uint64_t dah = umma_desc(a_hi + off), dal = umma_desc(a_lo + off);
            const uint64_t dbh = umma_desc(b_hi + off), dbl = umma_desc(b_lo + off);
            uint32_t accum = (kb | k) != 0;  // p.abl: precision ablation (common.cuh), 0 on the product path
            if (!(p.abl & 1)) { umma_tf32(d_tmem, dal, dbh, idesc, accum); accum = 1; }
            if (!(p.abl & 2)) { umma_tf32(d_tmem, dah, dbl, idesc, accum); accum = 1; }
            umma_tf32(d_tmem, dah, dbh, idesc, accum);
          }
          umma_commit(&empty[s]);          // frees the stage when these MMAs have read it
        }
        umma_commit(&tfull[acc]);          // accumulator complete
      }
    }
  } else if (warp >= SPLIT_WARP0) {
    // ------------------------------------------------------------------ splitters: lo = x - tf32(x)
    const int t = threadIdx.x - SPLIT_WARP0 * 32;  // 0..127
    uint32_t it = 0;
    auto lo4 = [](float4 v) {
      float4 r;
      r.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
      r.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
      r.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
      r.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
      return r;
    };
    if constexpr ((LNF & LN_A) != 0) {
      // A = LayerNorm(x) (LnEpi): thread t owns row t of the tile and rewrites it in place as
      // (x - mean) rstd gamma + beta before taking the lo part.  gamma / beta of the whole K range sit in shared
      // memory (zero beyond K, where TMA has zero-filled W as well); 16-byte chunk c of a 128-byte swizzle row is
      // stored at c ^ (row % 8), so the eight rows of a quarter-warp hit eight different bank groups.
      float* gb = reinterpret_cast<float*>(base + S * SM::STAGE + 256 + EPI_WARPS * 32 * CP * 4);  // [2][KB * BK]
      const int kpad = KB * BK;
      for (int i = t; i < kpad; i += 128) {
        gb[i] = i < p.K ? p.ln.a_g[i] : 0.f;
        gb[kpad + i] = i < p.K ? p.ln.a_b[i] : 0.f;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const uint32_t rowoff = (uint32_t)((t >> 3) * 1024 + (t & 7) * 128), sx = (uint32_t)(t & 7);
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int row = (tile / p.ntn) * BM + t;
        float rs = 1.f, nm = 0.f;
        if (row < p.M) {
          const float2 mr = p.ln.a_rs[row / p.ln.L];
          rs = mr.y;
          nm = -mr.x * mr.y;
        }
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const int s = it % S;
          mbar_wait(&full[s], (it / S) & 1);
          uint8_t* a = stA(s) + rowoff;
          uint8_t* al = stAlo(s) + rowoff;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint32_t o = ((uint32_t)c ^ sx) << 4;
            float4 v = *reinterpret_cast<const float4*>(a + o);
            const float4 g4 = *reinterpret_cast<const float4*>(gb + kb * BK + c * 4);
            const float4 e4 = *reinterpret_cast<const float4*>(gb + kpad + kb * BK + c * 4);
            v.x = fmaf(fmaf(v.x, rs, nm), g4.x, e4.x);
            v.y = fmaf(fmaf(v.y, rs, nm), g4.y, e4.y);
            v.z = fmaf(fmaf(v.z, rs, nm), g4.z, e4.z);
            v.w = fmaf(fmaf(v.w, rs, nm), g4.w, e4.w);
            *reinterpret_cast<float4*>(a + o) = v;
            *reinterpret_cast<float4*>(al + o) = lo4(v);
          }
          if (!p.has_blo) {
            const float4* b = reinterpret_cast<const float4*>(stB(s));
            float4* bl = reinterpret_cast<float4*>(stBlo(s));
#pragma unroll 4
            for (int i = t; i < SM::B_BYTES / 16; i += 128) bl[i] = lo4(b[i]);
          }
          fence_proxy_async();
          mbar_arrive(&split[s]);
        }
      }
    } else {
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      for (int kb = 0; kb < KB; ++kb, ++it) {
        const int s = it % S;
        long long t0 = clock64();
        mbar_wait(&full[s], (it / S) & 1);
        const long long t1 = clock64();
        dbg_acc[3] += t1 - t0;
        const float4* a = reinterpret_cast<const float4*>(stA(s));
        float4* al = reinterpret_cast<float4*>(stAlo(s));
        const float4* b = reinterpret_cast<const float4*>(stB(s));
        float4* bl = reinterpret_cast<float4*>(stBlo(s));
#pragma unroll 4
        for (int i = t; i < SM::A_BYTES / 16; i += 128) al[i] = lo4(a[i]);
        if (!p.has_blo) {
#pragma unroll 4
          for (int i = t; i < SM::B_BYTES / 16; i += 128) bl[i] = lo4(b[i]);
        }
        fence_proxy_async();  // generic-proxy writes -> visible to the tensor core (async proxy)
        mbar_arrive(&split[s]);
        dbg_acc[4] += clock64() - t1;
      }
    }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue
    // tcgen05.ld hands every thread one accumulator ROW (32 columns).  Writing rows straight to global
    // would cost 32 cache lines per instruction, so each 32x32 chunk is transposed through a padded
    // shared-memory tile and then handled 4 rows x 128 B per warp instruction: bias / periodic bias /
    // residual loads and the output store are all full-line, 16-byte-per-lane accesses, and every load
    // of a chunk is issued before the first store (C and R may alias as far as the compiler knows).
    const int q = warp & 3;            // TMEM lane quarter this warp may read
    const int half = (warp - 4) >> 2;  // the EPI_PARTS warps of a quarter take every EPI_PARTS-th 16-column sub-chunk
    float* stg = reinterpret_cast<float*>(base + S * SM::STAGE + 256) + (warp - 4) * (32 * CP);
    const bool all_al = (p.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) &&
                        (!p.R || ((p.ldr % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.R) & 15) == 0))) &&
                        (!p.bias || ((reinterpret_cast<uintptr_t>(p.bias) & 15) == 0)) &&
                        (!p.pbias || ((p.N % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.pbias) & 15) == 0)));
    constexpr int LPR = CW / 4;        // lanes per row (4)
    constexpr int RPI = 32 / LPR;      // rows per warp instruction (8)
    constexpr int NIT = 32 / RPI;      // iterations per sub-chunk (4)
    const int rsub = lane / LPR;       // this lane's row inside a group of RPI (staged path)
    uint32_t tcount = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tcount) {
      const int acc = tcount & 1;
      const int m0 = (tile / p.ntn) * BM, n0 = (tile % p.ntn) * BN;
      const int rbase = m0 + q * 32 + rsub;  // rows rbase + RPI*itr
      int pbo[NIT];                          // periodic-bias row offsets (row % period) * N
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
      int pbd[4] = {0, 0, 0, 0};  // direct path: periodic-bias offsets of rows m0 + q*32 + {0,8,16,24} + lane/4
      if (p.pbias && p.epi_direct) {
#pragma unroll
        for (int i = 0; i < 4; ++i) pbd[i] = ((m0 + q * 32 + i * 8 + (lane >> 2)) % p.pperiod) * p.N;
      }
      const long long t0 = clock64();
      mbar_wait(&tfull[acc], (tcount >> 1) & 1);
      const long long t1 = clock64();
      dbg_acc[5] += t1 - t0;
      tc_fence_after();
      float unused_cmax = 0.f;
      epilogue_tile<BN, ACT, false, (LNF & ~LN_A)>(p, tmem_base + acc * BN, m0, n0, q, half, lane, stg, all_al, pbo, pbd, 1.0f, unused_cmax);
      tc_fence_before();
      mbar_arrive(&tempty[acc]);
      dbg_acc[6] += clock64() - t1;
    }
  }
  if (p.dbg && lane == 0 && (warp == 0 || warp == 1 || warp == 4 || warp == 12)) {
    for (int i = 0; i < 7; ++i)
      if (dbg_acc[i]) p.dbg[blockIdx.x * 8 + i] = dbg_acc[i];
    if (warp == 0) p.dbg[blockIdx.x * 8 + 7] = clock64() - t_start;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * BN));
  }
}

}  // namespace tc

int ln_stats_parts(int N) { return tcc::EPI_PARTS * cdiv(N, N > 128 ? 256 : 128); }

// Can this GEMM go through the tensor-core kernel?  (W must be [N,K] K-major.)
bool gemm_tc_supported(const GemmArgs& a) {
  if (!a.w_is_nk) return false;
  if (a.M < 1 || a.K < 8) return false;
  if ((a.lda % 4) || (a.ldw % 4)) return false;  // TMA: 16-byte row pitch
  if ((reinterpret_cast<uintptr_t>(a.A) & 15) || (reinterpret_cast<uintptr_t>(a.W) & 15)) return false;
  return true;
}

int launch_gemm_tc(const GemmArgs& a, cudaStream_t st) {
  using namespace tc;
  static int num_sms = [] {
    int dev = 0, n = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n;
  }();
  if (a.M == 0) return 0;
  const int bn = a.N > 128 ? 256 : 128;
  CUtensorMap mapA, mapB, mapBlo;
  EDGL_TRY(make_map(&mapA, a.A, false, a.M, a.K, a.lda, BM));
  EDGL_TRY(make_map(&mapB, a.W, false, a.N, a.K, a.ldw, bn));
  static const bool no_blo = getenv("EDGL_TC_NOBLO") != nullptr;  // A/B switch for measurements
  const bool has_blo = a.Wlo != nullptr && !no_blo && (reinterpret_cast<uintptr_t>(a.Wlo) & 15) == 0;
  EDGL_TRY(make_map(&mapBlo, has_blo ? a.Wlo : a.W, false, a.N, a.K, a.ldw, bn));
  Params p;
  p.C = a.C; p.ldc = a.ldc; p.M = a.M; p.N = a.N; p.K = a.K; p.bias = a.bias; p.pbias = a.pbias;
  p.pperiod = a.pperiod > 0 ? a.pperiod : 1; p.R = a.R; p.ldr = a.ldr; p.act = a.act;
  p.col0_bias_only = a.zero_wrow0 ? 1 : 0;
  static const bool gelu_exact = [] { const char* e = getenv("EDGL_GELU"); return e && e[0] == 'e'; }();
  p.gelu_fit = gelu_exact ? 0 : 1;
  p.has_blo = has_blo ? 1 : 0;
  p.abl = ablation_gemm_bits();
  // Epilogue choice.  "direct" (TMEM fragments -> 32-byte sector stores, no shared-memory transpose) relieves the
  // shared-memory port - the resource these GEMMs are bound by - and measures 3-6 % faster when the epilogue has no
  // residual / periodic-bias rows to fetch (FF1, transform, logits); with them the 8-byte loads make it LSU-bound
  // and the staged 16-byte path wins (QKVT, attention-out, FF2).  EDGL_TC_EPI=direct|staged forces one.
  static const char epi_mode = [] {
    const char* e = getenv("EDGL_TC_EPI");
    return e ? e[0] : 'a';
  }();
  p.epi_direct = epi_mode == 'd' ? 1 : epi_mode == 's' ? 0 : (a.R == nullptr && a.pbias == nullptr) ? 1 : 0;
  p.ln = a.ln;
  p.flt = a.flt;
  p.run_if = a.run_if;
  int lnf = 0;
  if (a.ln.any()) {
    // the fused LayerNorm pieces live in the staged fast path of the epilogue (and in the splitter warps) only
    EDGL_REQUIRE(a.N % 16 == 0 && a.ldc % 4 == 0 && (!a.R || a.ldr % 4 == 0) && a.ln.L >= 1 && (!a.ln.a_rs || a.K <= 256) &&
                     (reinterpret_cast<uintptr_t>(a.C) & 15) == 0 && (!a.R || (reinterpret_cast<uintptr_t>(a.R) & 15) == 0) &&
                     (!a.bias || (reinterpret_cast<uintptr_t>(a.bias) & 15) == 0) && !a.pbias,
                 "gemm_tc: fused LayerNorm needs N %% 16 == 0, K <= 256 for a normalised A, and 16-byte aligned operands");
    EDGL_REQUIRE(!a.ln.a_rs || (a.ln.a_g && a.ln.a_b), "gemm_tc: LayerNorm on A needs gamma and beta");
    EDGL_REQUIRE(!a.ln.r_rs || (a.R && a.ln.r_g && a.ln.r_b), "gemm_tc: LayerNorm on the residual needs R, gamma, beta");
    p.epi_direct = 0;
    lnf = (a.ln.a_rs ? LN_A : 0) | (a.ln.r_rs ? LN_RES : 0) | (a.ln.stats ? LN_STATS : 0) | (a.ln.last_only ? LN_LAST : 0);
  }
  p.ntn = cdiv(a.N, bn);
  const long long ntm = cdiv(a.M, BM);
  EDGL_REQUIRE(ntm * p.ntn < (1ll << 31), "gemm_tc: too many tiles");
  p.num_tiles = (int)(ntm * p.ntn);
  p.kblocks = cdiv(a.K, BK);
  const int grid = p.num_tiles < num_sms ? p.num_tiles : num_sms;
  p.dbg = nullptr;
  static const bool debug = getenv("EDGL_TC_DEBUG") != nullptr;
  static long long* dbg_buf = nullptr;
  if (debug) {
    if (!dbg_buf) cudaMalloc(&dbg_buf, 8 * 256 * sizeof(long long));
    cudaMemsetAsync(dbg_buf, 0, 8 * 256 * sizeof(long long), st);
    p.dbg = dbg_buf;
  }
#define EDGL_TC_LAUNCH_L(BNV, ACTV, LNV)                                                                       \
  {                                                                                                            \
    auto kern = gemm_tc_kernel<BNV, ACTV, LNV>;                                                                \
    EDGL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<BNV>::BYTES));      \
    kern<<<grid, NTHREADS, Smem<BNV>::BYTES, st>>>(mapA, mapB, mapBlo, p);                                     \
  }
  if (a.flt.cand) {  // candidate filter of the logits layer: bias only
    EDGL_REQUIRE(lnf == 0 && a.act == ACT_NONE && !a.R && !a.pbias && a.flt.thr && a.flt.cnt && a.flt.cap > 0,
                 "gemm_tc: the top-K filter epilogue takes a plain bias layer");
    if (bn == 256) EDGL_TC_LAUNCH_L(256, ACT_NONE, LN_FILT)
    else EDGL_TC_LAUNCH_L(128, ACT_NONE, LN_FILT)
    EDGL_LAUNCH_CHECK();
    return 0;
  }
  // the fused-LayerNorm combinations the EasyDGL block tail uses (api.cu): attention-out (statistics), FF1 (A is a
  // LayerNorm; GELU), FF2 (residual is a LayerNorm; statistics), transform (A is a LayerNorm; GELU; statistics; last rows)
  if (lnf != 0) {
    bool ok = true;
    if (bn == 256) {
      if (lnf == LN_STATS && a.act == ACT_NONE) EDGL_TC_LAUNCH_L(256, ACT_NONE, LN_STATS)
      else if (lnf == LN_A && a.act == ACT_GELU) EDGL_TC_LAUNCH_L(256, ACT_GELU, LN_A)
      else if (lnf == (LN_RES | LN_STATS) && a.act == ACT_NONE) EDGL_TC_LAUNCH_L(256, ACT_NONE, LN_RES | LN_STATS)
      else if (lnf == (LN_A | LN_STATS | LN_LAST) && a.act == ACT_GELU) EDGL_TC_LAUNCH_L(256, ACT_GELU, LN_A | LN_STATS | LN_LAST)
      else ok = false;
    } else {
      if (lnf == LN_STATS && a.act == ACT_NONE) EDGL_TC_LAUNCH_L(128, ACT_NONE, LN_STATS)
      else if (lnf == LN_A && a.act == ACT_GELU) EDGL_TC_LAUNCH_L(128, ACT_GELU, LN_A)
      else if (lnf == (LN_RES | LN_STATS) && a.act == ACT_NONE) EDGL_TC_LAUNCH_L(128, ACT_NONE, LN_RES | LN_STATS)
      else if (lnf == (LN_A | LN_STATS | LN_LAST) && a.act == ACT_GELU) EDGL_TC_LAUNCH_L(128, ACT_GELU, LN_A | LN_STATS | LN_LAST)
      else ok = false;
    }
    EDGL_REQUIRE(ok, "gemm_tc: fused-LayerNorm combination %d with activation %d is not instantiated", lnf, a.act);
    EDGL_LAUNCH_CHECK();
    return 0;
  }
#define EDGL_TC_LAUNCH(BNV, ACTV)                                                                              \
  {                                                                                                            \
    auto kern = gemm_tc_kernel<BNV, ACTV, 0>;                                                                  \
    EDGL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<BNV>::BYTES));      \
    kern<<<grid, NTHREADS, Smem<BNV>::BYTES, st>>>(mapA, mapB, mapBlo, p);                                            \
  }
  if (bn == 256) {
    if (a.act == ACT_GELU) EDGL_TC_LAUNCH(256, ACT_GELU)
    else if (a.act == ACT_RELU) EDGL_TC_LAUNCH(256, ACT_RELU)
    else EDGL_TC_LAUNCH(256, ACT_NONE)
  } else {
    if (a.act == ACT_GELU) EDGL_TC_LAUNCH(128, ACT_GELU)
    else if (a.act == ACT_RELU) EDGL_TC_LAUNCH(128, ACT_RELU)
    else EDGL_TC_LAUNCH(128, ACT_NONE)
  }
#undef EDGL_TC_LAUNCH
  EDGL_LAUNCH_CHECK();
  if (debug) {
    long long hbuf[8 * 256];
    cudaStreamSynchronize(st);
    cudaMemcpy(hbuf, dbg_buf, sizeof(hbuf), cudaMemcpyDeviceToHost);
    double s7[8] = {0};
    for (int b = 0; b < grid; ++b)
      for (int i = 0; i < 8; ++i) s7[i] += (double)hbuf[b * 8 + i] / grid;
    fprintf(stderr,
            "[tc] M=%d N=%d K=%d bn=%d tiles/cta=%.1f | cycles/CTA: total %.0f | producer wait_empty %.0f | mma wait_tempty "
            "%.0f wait_split %.0f | splitter wait_full %.0f work %.0f | epilogue wait_tfull %.0f work %.0f\n",
            a.M, a.N, a.K, bn, (double)p.num_tiles / grid, s7[7], s7[0], s7[1], s7[2], s7[3], s7[4], s7[5], s7[6]);
  }
  return 0;
}

}  // namespace edgl
