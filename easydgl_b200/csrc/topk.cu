// topk.cu - ranking part of Sequential.eval (Base.py:150-181).
//   mask_seen : logits[b, seqs_i[b,l]] = -inf for every l (Base.py:156-163; adding -inf to a finite logit)
//   topk      : tf.nn.top_k(., k) - sorted descending, ties -> LOWER index first (Base.py:181)
//   merge     : K-way merge of per-shard candidate lists (multi-GPU)
// Ranking is done on the masked logits: softmax (Base.py:164) is monotone, so the order is the same
// wherever fp32 softmax is injective (DESIGN.md discusses the underflow corner).
// Integer/index work: bit-exact by construction (order-preserving uint keys, radix select, ties by index).
#include <stdlib.h>
#include <string.h>

#include <map>
#include <mutex>
#include <utility>

#include "common.cuh"

namespace edgl {

// f2key / key2f / compose: common.cuh (shared with the candidate-filter epilogue of the logits GEMM)

__global__ void mask_seen_kernel(float* __restrict__ logits, int ld, long long n, int seen_len, long long seen_stride,
                                 const int64_t* __restrict__ ids, long long col0, long long col1,
                                 const int* __restrict__ run_if) {
  if (run_if && *run_if == 0) return;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long b = i / seen_len;
  const long long id = ids[b * seen_stride + (i - b * seen_len)];
  if (id >= col0 && id < col1) logits[b * ld + (id - col0)] = -INFINITY;
}

int launch_mask_seen(float* logits, int ld, int B, const int64_t* ids, int seen_len, long long seen_stride,
                     long long col0, long long col1, cudaStream_t st, const int* run_if) {
  const long long n = (long long)B * seen_len;
  if (n == 0) return 0;
  mask_seen_kernel<<<cdiv(n, 256), 256, 0, st>>>(logits, ld, n, seen_len, seen_stride, ids, col0, col1, run_if);
  EDGL_LAUNCH_CHECK();
  return 0;
}

// descending bitonic sort of n (power of two) u64 in shared memory, blockDim.x threads
__device__ void bitonic_desc(unsigned long long* s, int n) {
  for (int k = 2; k <= n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      __syncthreads();
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long a = s[i], b = s[ixj];
          const bool desc = (i & k) == 0;
          if (desc ? (a < b) : (a > b)) {
            s[i] = b;
            s[ixj] = a;
          }
        }
      }
    }
  }
  __syncthreads();
}

// ---- warp-shuffle stages of the same bitonic network (element index i, stage k, distance j <= 16)
__device__ __forceinline__ unsigned long long cmpx64(unsigned long long v, int i, int k, int j) {
  const unsigned long long o = __shfl_xor_sync(0xffffffffu, v, j);
  const bool take_max = (((i & j) == 0) == ((i & k) == 0));  // lower partner of a descending block keeps the max
  return take_max ? (v > o ? v : o) : (v < o ? v : o);
}
__device__ __forceinline__ uint32_t cmpx32(uint32_t v, int i, int k, int j) {
  const uint32_t o = __shfl_xor_sync(0xffffffffu, v, j);
  const bool take_max = (((i & j) == 0) == ((i & k) == 0));
  return take_max ? max(v, o) : min(v, o);
}

// Descending bitonic sort of n (power of two, >= 32) u64 in shared memory by 256 threads: every
// compare-exchange at distance <= 16 runs in registers with warp shuffles (no barrier); only the
// log2(n)-5 long-distance steps of the last stages touch shared memory between barriers.
__device__ void bitonic_desc_hybrid(unsigned long long* s, int n) {
  const int lane = threadIdx.x & 31;
  __syncthreads();
  for (int base = (threadIdx.x >> 5) * 32; base < n; base += 256) {
    const int i = base + lane;
    unsigned long long v = s[i];
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1)
#pragma unroll
      for (int j = k >> 1; j > 0; j >>= 1) v = cmpx64(v, i, k, j);
    s[i] = v;
  }
  for (int k = 64; k <= n; k <<= 1) {
    for (int j = k >> 1; j >= 32; j >>= 1) {
      __syncthreads();
      for (int i = threadIdx.x; i < n; i += 256) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long a = s[i], b = s[ixj];
          if (((i & k) == 0) ? (a < b) : (a > b)) {
            s[i] = b;
            s[ixj] = a;
          }
        }
      }
    }
    __syncthreads();
    for (int base = (threadIdx.x >> 5) * 32; base < n; base += 256) {
      const int i = base + lane;
      unsigned long long v = s[i];
#pragma unroll
      for (int j = 16; j > 0; j >>= 1) v = cmpx64(v, i, k, j);
      s[i] = v;
    }
  }
  __syncthreads();
}

// first K entries of the sorted candidate list -> (global index, value); pad when fewer than K exist
__device__ __forceinline__ void write_topk(const unsigned long long* cand, int Keff, int K, int col_offset,
                                           int32_t* __restrict__ idx_out, float* __restrict__ val_out) {
  for (int i = threadIdx.x; i < K; i += blockDim.x) {
    if (i < Keff) {
      const unsigned long long c = cand[i];
      idx_out[i] = (int32_t)(0xffffffffu - (uint32_t)(c & 0xffffffffull)) + col_offset;
      val_out[i] = key2f((uint32_t)(c >> 32));
    } else {
      idx_out[i] = -1;
      val_out[i] = -INFINITY;
    }
  }
}

// Exact radix select of one row (any N, any tie pattern): 4 passes of 8-bit histograms over the
// order-preserving keys, then an index-ordered pass for the ties at the K-th value, then a bitonic sort.
// One CTA (256 threads); cand = KP u64 in shared memory.
__device__ void topk_row_radix(const float* __restrict__ p, int N, int K, int KP, int col_offset,
                               unsigned long long* cand, int32_t* __restrict__ idx_out, float* __restrict__ val_out) {
  __shared__ unsigned int hist[256];
  __shared__ unsigned int s_prefix, s_remaining, s_eq_total, s_gt_cnt;
  __shared__ unsigned int warp_tot[8];
  const int tid = threadIdx.x;
  const int Keff = K < N ? K : N;
  __syncthreads();

  if (tid == 0) {
    s_prefix = 0;
    s_remaining = Keff;
    s_gt_cnt = 0;
  }
  for (int i = tid; i < KP; i += 256) cand[i] = 0ull;
  uint32_t mask = 0;
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    hist[tid] = 0;
    __syncthreads();
    const uint32_t prefix = s_prefix;
    for (int i = tid; i < N; i += 256) {
      const uint32_t key = f2key(p[i]);
      if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (tid == 0) {
      unsigned int cum = 0, rem = s_remaining;
      int dgt = 255;
      for (; dgt > 0; --dgt) {
        if (cum + hist[dgt] >= rem) break;
        cum += hist[dgt];
      }
      s_remaining = rem - cum;       // how many still to take among keys with this digit
      s_eq_total = hist[dgt];
      s_prefix = prefix | ((uint32_t)dgt << shift);
    }
    mask |= 0xffu << shift;
    __syncthreads();
  }
  const uint32_t T = s_prefix;              // the Keff-th largest key
  const unsigned int need_eq = s_remaining; // ties at T to keep (lowest indices first)
  const unsigned int n_gt = Keff - need_eq;
  const bool all_eq = (s_eq_total == need_eq);
  __syncthreads();

  if (all_eq) {
    for (int i = tid; i < N; i += 256) {
      const uint32_t key = f2key(p[i]);
      if (key >= T) {
        const unsigned int pos = atomicAdd(&s_gt_cnt, 1u);
        cand[pos] = compose(key, (uint32_t)i);
      }
    }
  } else {
    // ordered pass: ties must be taken in index order
    unsigned int eq_taken = 0;
    for (int base = 0; base < N && eq_taken < need_eq; base += 256) {
      const int i = base + tid;
      uint32_t key = 0;
      if (i < N) key = f2key(p[i]);
      const bool is_eq = (i < N) && key == T;
      const unsigned int bal = __ballot_sync(0xffffffffu, is_eq);
      const int lane = tid & 31, w = tid >> 5;
      if (lane == 0) warp_tot[w] = __popc(bal);
      __syncthreads();
      unsigned int before = 0, total = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (j < w) before += warp_tot[j];
        total += warp_tot[j];
      }
      if (is_eq) {
        const unsigned int pos = eq_taken + before + __popc(bal & ((1u << lane) - 1u));
        if (pos < need_eq) cand[n_gt + pos] = compose(key, (uint32_t)i);
      }
      eq_taken += total;
      __syncthreads();
    }
    for (int i = tid; i < N; i += 256) {
      const uint32_t key = f2key(p[i]);
      if (key > T) {
        const unsigned int pos = atomicAdd(&s_gt_cnt, 1u);
        cand[pos] = compose(key, (uint32_t)i);
      }
    }
  }
  bitonic_desc(cand, KP);
  write_topk(cand, Keff, K, col_offset, idx_out, val_out);
}

// Fast path for K <= 256, N >= 1024: every thread takes the maximum of its strided slice; a threshold t0
// derived from the slice maxima (see below) is a LOWER bound of the row's K-th largest value (at least K
// elements are >= t0), so {x >= t0} contains the whole top-K including every tie at the cut.  Typically
// ~1.9 K candidates survive; they are sorted exactly (key desc, index asc).  Rows that overflow the candidate
// buffer (constant rows, everything masked) take the radix path.  Two coalesced reads of the row.
constexpr int kCandCap = 1024;

// system-scope release store / acquire load for the cross-GPU flags (peer memory over NVLink)
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Called by every thread of every CTA after its last peer write: the last CTA of the grid publishes
// `epoch` into slot `my_rank` of every peer's flag array (the writes of all CTAs are fenced before it).
__device__ void p2p_signal_when_grid_done(unsigned int* counter, const long long* peer_flags, int G, int my_rank,
                                          uint32_t epoch) {
  __syncthreads();  // every thread's peer stores are ordered before thread 0 ...
  if (threadIdx.x == 0) {
    __threadfence_system();  // ... whose system-scope fence is cumulative over them
    const unsigned int prev = atomicAdd(counter, 1u);
    if (prev == gridDim.x - 1) {
      *counter = 0;
      __threadfence_system();
      for (int g = 0; g < G; ++g) st_release_sys(reinterpret_cast<uint32_t*>(peer_flags[g]) + my_rank, epoch);
    }
  }
}

__device__ void topk_row(const float* __restrict__ p, int ld_vec_ok, int N, int K, int KP, int col_offset,
                         unsigned long long* cand, int32_t* __restrict__ io, float* __restrict__ vo);

__global__ void __launch_bounds__(256) topk_kernel(const float* __restrict__ logits, int ld, int N, int K, int KP,
                                                   int col_offset, long long out_stride,
                                                   int32_t* __restrict__ idx_out, float* __restrict__ val_out,
                                                   TopkP2P pp, const int* __restrict__ run_if) {
  extern __shared__ __align__(16) unsigned long long cand[];  // max(KP, kCandCap) entries
  if (run_if && *run_if == 0) return;
  const float* p = logits + (long long)blockIdx.x * ld;
  int32_t* io;
  float* vo;
  if (pp.dest) {
    // fused exchange: row R of the gathered batch belongs to rank R / rows_per_dest; its candidates go
    // straight into that rank's receive buffer (block my_rank, interleaved [idx | val]) over NVLink
    const long long R = (long long)pp.row_base + blockIdx.x;
    const int dst = (int)(R / pp.rows_per_dest);
    io = reinterpret_cast<int32_t*>(pp.dest[dst]) +
         ((long long)pp.my_rank * pp.rows_per_dest + (R - (long long)dst * pp.rows_per_dest)) * (2 * K);
    vo = reinterpret_cast<float*>(io + K);
  } else {
    io = idx_out + (long long)blockIdx.x * out_stride;
    vo = val_out + (long long)blockIdx.x * out_stride;
  }
  const int vec = (ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(logits) & 15) == 0);
  topk_row(p, vec, N, K, KP, col_offset, cand, io, vo);
  if (pp.dest && pp.signal) p2p_signal_when_grid_done(pp.counter, pp.peer_flags, pp.G, pp.my_rank, pp.epoch);
}

__device__ void topk_row(const float* __restrict__ p, int vec, int N, int K, int KP, int col_offset,
                         unsigned long long* cand, int32_t* __restrict__ io, float* __restrict__ vo) {
  __shared__ unsigned int s_cnt;
  __shared__ uint32_t s_t0, s_wt[8];
  const int tid = threadIdx.x;
  if (K > 256 || N < 1024) {
    topk_row_radix(p, N, K, KP, col_offset, cand, io, vo);
    return;
  }
  const int N4 = vec ? (N >> 2) : 0;
  // ---- pass 1: slice maxima
  uint32_t mx = 0;
  for (int i = tid; i < N4; i += 256) {
    const float4 v = reinterpret_cast<const float4*>(p)[i];
    mx = max(max(mx, f2key(v.x)), max(f2key(v.y), max(f2key(v.z), f2key(v.w))));
  }
  for (int i = 4 * N4 + tid; i < N; i += 256) mx = max(mx, f2key(p[i]));
  // per-warp: the q-th largest of the 32 slice maxima, q = ceil(K/8); the minimum of those over the 8
  // warps has at least 8*q >= K elements at or above it -> a valid lower bound of the K-th largest value
  {
    const int lane = tid & 31;
    uint32_t v = mx;
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1)
#pragma unroll
      for (int j = k >> 1; j > 0; j >>= 1) v = cmpx32(v, lane, k, j);  // lanes < 32: the last stage is descending
    const int q = (K + 7) >> 3;
    if (lane == q - 1) s_wt[tid >> 5] = v;
    if (tid == 0) s_cnt = 0;
    __syncthreads();
    uint32_t t = s_wt[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) t = min(t, s_wt[w]);
    if (tid == 0) s_t0 = t;
  }
  __syncthreads();
  const uint32_t t0 = s_t0;
  // ---- pass 2: collect {key >= t0}
  for (int i = tid; i < N4; i += 256) {
    const float4 v = reinterpret_cast<const float4*>(p)[i];
    const uint32_t k4[4] = {f2key(v.x), f2key(v.y), f2key(v.z), f2key(v.w)};
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (k4[e] >= t0) {
        const unsigned int pos = atomicAdd(&s_cnt, 1u);
        if (pos < kCandCap) cand[pos] = compose(k4[e], (uint32_t)(4 * i + e));
      }
  }
  for (int i = 4 * N4 + tid; i < N; i += 256) {
    const uint32_t k = f2key(p[i]);
    if (k >= t0) {
      const unsigned int pos = atomicAdd(&s_cnt, 1u);
      if (pos < kCandCap) cand[pos] = compose(k, (uint32_t)i);
    }
  }
  __syncthreads();
  const unsigned int cnt = s_cnt;
  if (cnt > kCandCap) {  // block-uniform
    topk_row_radix(p, N, K, KP, col_offset, cand, io, vo);
    return;
  }
  int np = 128;
  while (np < (int)cnt) np <<= 1;
  for (int i = cnt + tid; i < np; i += 256) cand[i] = 0ull;
  bitonic_desc_hybrid(cand, np);
  write_topk(cand, K, K, col_offset, io, vo);
}

// ---- short rows (multi-GPU column shards: N of a few thousand, G * B rows): ONE WARP PER ROW, no block barrier.
// Every lane keeps N / 32 order-preserving keys in registers; the K-th largest key is found by a bit-wise descent
// (32 counting passes over the registers), ties at the cut are taken in index order with ballots, and the K selected
// (key, index) pairs are sorted by a 128-element bitonic network that lives in the warp's registers.  Same result,
// bit for bit, as topk_row: (value desc, index asc).  The CTA-per-row kernel spends ~14 us of barriers and shared-
// memory sorting per row whatever its length, which is what limited the 8-GPU step (32768 rows of 2251 columns).
template <int NPL>
__global__ void __launch_bounds__(256) topk_warp_kernel(const float* __restrict__ logits, int ld, long long rows, int N,
                                                        int K, int col_offset, long long out_stride,
                                                        int32_t* __restrict__ idx_out, float* __restrict__ val_out,
                                                        TopkP2P pp) {
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + w;
  if (row < rows) {  // warp-uniform
    const float* p = logits + row * ld;
    uint32_t k[NPL];
#pragma unroll
    for (int j = 0; j < NPL; ++j) {
      const int i = j * 32 + lane;
      k[j] = i < N ? f2key(p[i]) : 0u;  // 0 is below the key of every float
    }
    const int Keff = K < N ? K : N;
    uint32_t T = 0;  // becomes the Keff-th largest key: the largest T with #{key >= T} >= Keff
#pragma unroll 1
    for (int bit = 31; bit >= 0; --bit) {
      const uint32_t c = T | (1u << bit);
      int cnt = 0;
#pragma unroll
      for (int j = 0; j < NPL; ++j) cnt += (k[j] >= c) ? 1 : 0;
      cnt = __reduce_add_sync(0xffffffffu, cnt);
      if (cnt >= Keff) T = c;
    }
    int ngt = 0;
#pragma unroll
    for (int j = 0; j < NPL; ++j) ngt += (k[j] > T) ? 1 : 0;
    ngt = __reduce_add_sync(0xffffffffu, ngt);
    const int need_eq = Keff - ngt;  // ties at the cut, lowest indices first
    // the selected pairs, element e of the sort network lives in v[e / 32] of lane e % 32
    unsigned long long v[4] = {0ull, 0ull, 0ull, 0ull};
    int base_gt = 0, base_eq = 0;
    const unsigned int lt = (1u << lane) - 1u;
#pragma unroll
    for (int j = 0; j < NPL; ++j) {
      const int i = j * 32 + lane;
      const bool gt = k[j] > T, eq = (k[j] == T) && (i < N);
      const unsigned int bg = __ballot_sync(0xffffffffu, gt), be = __ballot_sync(0xffffffffu, eq);
      int pos = -1;
      if (gt) pos = base_gt + __popc(bg & lt);
      if (eq) {
        const int pe = base_eq + __popc(be & lt);
        if (pe < need_eq) pos = ngt + pe;
      }
      base_gt += __popc(bg);
      base_eq += __popc(be);
      // hand the pair to the lane / register that owns sort slot `pos` (at most 32 selected per j)
      const unsigned long long mine = compose(k[j], (uint32_t)i);
      unsigned int sel = __ballot_sync(0xffffffffu, pos >= 0);
      while (sel) {
        const int src = __ffs(sel) - 1;
        sel &= sel - 1;
        const int ps = __shfl_sync(0xffffffffu, pos, src);
        const unsigned long long val = __shfl_sync(0xffffffffu, mine, src);
        if ((ps & 31) == lane) {
          if ((ps >> 5) == 0) v[0] = val;
          else if ((ps >> 5) == 1) v[1] = val;
          else if ((ps >> 5) == 2) v[2] = val;
          else v[3] = val;
        }
      }
    }
    // descending bitonic sort of the 128 slots
#pragma unroll
    for (int kk = 2; kk <= 128; kk <<= 1) {
#pragma unroll
      for (int j = kk >> 1; j > 0; j >>= 1) {
        if (j >= 32) {
          const int dm = j >> 5;
#pragma unroll
          for (int m = 0; m < 4; ++m) {
            if ((m & dm) == 0) {
              const int e = m * 32 + lane;
              const bool desc = (e & kk) == 0;
              const unsigned long long a = v[m], b = v[m | dm];
              const bool sw = desc ? (a < b) : (a > b);
              v[m] = sw ? b : a;
              v[m | dm] = sw ? a : b;
            }
          }
        } else {
#pragma unroll
          for (int m = 0; m < 4; ++m) v[m] = cmpx64(v[m], m * 32 + lane, kk, j);
        }
      }
    }
    int32_t* io;
    float* vo;
    if (pp.dest) {
      const long long R = (long long)pp.row_base + row;
      const int dst = (int)(R / pp.rows_per_dest);
      io = reinterpret_cast<int32_t*>(pp.dest[dst]) +
           ((long long)pp.my_rank * pp.rows_per_dest + (R - (long long)dst * pp.rows_per_dest)) * (2 * K);
      vo = reinterpret_cast<float*>(io + K);
    } else {
      io = idx_out + row * out_stride;
      vo = val_out + row * out_stride;
    }
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const int e = m * 32 + lane;
      if (e < K) {
        if (e < Keff) {
          io[e] = (int32_t)(0xffffffffu - (uint32_t)(v[m] & 0xffffffffull)) + col_offset;
          vo[e] = key2f((uint32_t)(v[m] >> 32));
        } else {
          io[e] = -1;
          vo[e] = -INFINITY;
        }
      }
    }
  }
  if (pp.dest && pp.signal) p2p_signal_when_grid_done(pp.counter, pp.peer_flags, pp.G, pp.my_rank, pp.epoch);
}

// ---- short rows, second version (default for the column shards of the multi-GPU path): one warp per row, ~40
// registers, two passes over the row instead of a register-resident copy.  Pass 1 takes 512 GROUP maxima (16 strided
// groups per lane); the Keff-th largest of them, resolved to its 16 high bits, is a lower bound T0 of the row's Keff-th
// largest key (at least Keff elements are >= T0).  Pass 2 appends every key >= T0 - typically ~110 of 2251, all ties
// at the cut included - to a per-warp list in shared memory; the list is sorted by the 128-slot register network
// (composed (key, ~index) words: value descending, index ascending) and its first Keff entries are the row's top-K,
// bit for bit what topk_row returns.  Rows with more than 128 candidates (heavy ties, constant rows) are appended to
// an overflow list that topk_list_kernel works off with the CTA-per-row code.
template <bool VEC>  // VEC: 16-byte row pitch and base: the row is read as float4 (a lane owns 4 consecutive columns)
__global__ void __launch_bounds__(256) topk_warp2_kernel(const float* __restrict__ logits, int ld, long long rows, int N,
                                                         int K, int col_offset, long long out_stride,
                                                         int32_t* __restrict__ idx_out, float* __restrict__ val_out,
                                                         int* __restrict__ ovf_rows, unsigned int* __restrict__ ovf_cnt) {
  __shared__ unsigned long long s_list[8][128];
  __shared__ unsigned int s_cnt[8];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + w;
  if (row >= rows) return;  // warp-uniform; no block-wide barrier below
  const float* p = logits + row * ld;
  const int Keff = K < N ? K : N;
  const int npl = (N + 31) >> 5;
  const int n4 = (N + 3) >> 2, nit = (n4 + 31) >> 5;  // VEC: float4 words per row, iterations of a warp over them
  const float4* p4 = reinterpret_cast<const float4*>(p);
  uint32_t gm[16];
#pragma unroll
  for (int g = 0; g < 16; ++g) gm[g] = 0u;
  if constexpr (VEC) {
    for (int j0 = 0; j0 < nit; j0 += 16) {
#pragma unroll
      for (int g = 0; g < 16; ++g) {
        const int q = (j0 + g) * 32 + lane;
        if (q < n4) {
          const float4 x = p4[q];
          const int i = 4 * q;  // columns i .. i+3; those >= N (row padding) count as key 0, below every float
          const uint32_t k0 = f2key(x.x), k1 = i + 1 < N ? f2key(x.y) : 0u, k2 = i + 2 < N ? f2key(x.z) : 0u,
                         k3 = i + 3 < N ? f2key(x.w) : 0u;
          gm[g] = max(max(gm[g], k0), max(k1, max(k2, k3)));
        }
      }
    }
  } else {
    for (int j0 = 0; j0 < npl; j0 += 16) {
#pragma unroll
      for (int g = 0; g < 16; ++g) {
        const int i = (j0 + g) * 32 + lane;
        const uint32_t key = i < N ? f2key(p[i]) : 0u;  // 0 is below the key of every float
        gm[g] = max(gm[g], key);
      }
    }
  }
  uint32_t T0 = 0;
#pragma unroll 1
  for (int bit = 31; bit >= 16; --bit) {
    const uint32_t c = T0 | (1u << bit);
    int cnt = 0;
#pragma unroll
    for (int g = 0; g < 16; ++g) cnt += (gm[g] >= c) ? 1 : 0;
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if (cnt >= Keff) T0 = c;
  }
  if (lane == 0) s_cnt[w] = 0u;
  __syncwarp();
  auto take = [&](uint32_t key, int i) {
    if (i < N && key >= T0) {
      const unsigned int pos = atomicAdd(&s_cnt[w], 1u);
      if (pos < 128u) s_list[w][pos] = compose(key, (uint32_t)i);
    }
  };
  if constexpr (VEC) {
#pragma unroll 4
    for (int j = 0; j < nit; ++j) {
      const int q = j * 32 + lane;
      if (q < n4) {
        const float4 x = p4[q];
        take(f2key(x.x), 4 * q);
        take(f2key(x.y), 4 * q + 1);
        take(f2key(x.z), 4 * q + 2);
        take(f2key(x.w), 4 * q + 3);
      }
    }
  } else {
#pragma unroll 8
    for (int j = 0; j < npl; ++j) {
      const int i = j * 32 + lane;
      if (i < N) take(f2key(p[i]), i);
    }
  }
  __syncwarp();
  const unsigned int total = s_cnt[w];
  if (total > 128u) {  // warp-uniform
    if (lane == 0) ovf_rows[atomicAdd(ovf_cnt, 1u)] = (int)row;
    return;
  }
  unsigned long long v[4];  // sort slot e lives in v[e / 32] of lane e % 32
#pragma unroll
  for (int m = 0; m < 4; ++m) v[m] = (unsigned int)(m * 32 + lane) < total ? s_list[w][m * 32 + lane] : 0ull;
#pragma unroll
  for (int kk = 2; kk <= 128; kk <<= 1) {
#pragma unroll
    for (int j = kk >> 1; j > 0; j >>= 1) {
      if (j >= 32) {
        const int dm = j >> 5;
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          if ((m & dm) == 0) {
            const int e = m * 32 + lane;
            const bool desc = (e & kk) == 0;
            const unsigned long long a = v[m], b = v[m | dm];
            const bool sw = desc ? (a < b) : (a > b);
            v[m] = sw ? b : a;
            v[m | dm] = sw ? a : b;
          }
        }
      } else {
#pragma unroll
        for (int m = 0; m < 4; ++m) v[m] = cmpx64(v[m], m * 32 + lane, kk, j);
      }
    }
  }
  int32_t* io = idx_out + row * out_stride;
  float* vo = val_out + row * out_stride;
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    const int e = m * 32 + lane;
    if (e < K) {
      if (e < Keff) {
        io[e] = (int32_t)(0xffffffffu - (uint32_t)(v[m] & 0xffffffffull)) + col_offset;
        vo[e] = key2f((uint32_t)(v[m] >> 32));
      } else {
        io[e] = -1;
        vo[e] = -INFINITY;
      }
    }
  }
}

// the rows topk_warp2_kernel could not finish: CTA-per-row code over the overflow list (a fixed, small grid)
__global__ void __launch_bounds__(256) topk_list_kernel(const float* __restrict__ logits, int ld, int N, int K, int KP,
                                                        int col_offset, long long out_stride,
                                                        int32_t* __restrict__ idx_out, float* __restrict__ val_out,
                                                        const int* __restrict__ ovf_rows,
                                                        const unsigned int* __restrict__ ovf_cnt) {
  extern __shared__ __align__(16) unsigned long long cand[];  // max(KP, kCandCap) entries
  const unsigned int n = *ovf_cnt;
  const int vec = (ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(logits) & 15) == 0);
  for (unsigned int r = blockIdx.x; r < n; r += gridDim.x) {
    const long long row = ovf_rows[r];
    topk_row(logits + row * ld, vec, N, K, KP, col_offset, cand, idx_out + row * out_stride, val_out + row * out_stride);
    __syncthreads();
  }
}

static int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

int launch_topk(const float* logits, int ld, int B, int N, int K, int col_offset, long long out_stride, int32_t* idx,
                float* val, cudaStream_t st, const TopkP2P* p2p, const int* run_if) {
  if (out_stride == 0) out_stride = K;
  TopkP2P pp;
  memset(&pp, 0, sizeof(pp));
  if (p2p) pp = *p2p;
  EDGL_REQUIRE(K >= 1 && K <= 2048, "topk: K must be in [1,2048] (got %d)", K);
  EDGL_REQUIRE(N >= 1, "topk: N must be >= 1");
  if (B == 0) return 0;
  // short rows, many of them (column shards of the multi-GPU path): one warp per row.  Opt-in (EDGL_TOPK_WARP=1,
  // read per call): bit-identical, but measured SLOWER than the CTA kernel at the shape it was written for (8 GPUs,
  // 32768 rows x 2251 columns: 0.65 ms vs 0.39 ms) - the 32 counting passes over 72 registers cost ~7 K
  // instructions per row.
  const char* we = getenv("EDGL_TOPK_WARP");
  const bool use_warp = we != nullptr && we[0] == '1';
  // second warp-per-row version: the default for short rows (EDGL_TOPK_WARP=0 keeps the CTA kernel, =1 the first version)
  // measured against the CTA kernel (tools/bench_topk_rows.py): 32768 x 2252: 0.176 vs 0.387 ms; 16384 x 6252: 0.200 vs
  // 0.256; 8192 x 12504: 0.186 vs 0.173; 4096 x 18004: 0.201 vs 0.114 - so rows of up to 8192 columns take it
  static const int warp_max_n = [] { const char* e = getenv("EDGL_TOPK_WARP_MAXN"); return e ? atoi(e) : 8192; }();
  if (!(we && (we[0] == '0' || we[0] == '1')) && !p2p && !run_if && K <= 128 && N >= 256 && N <= warp_max_n && B >= 64) {
    // overflow list: one buffer per (device, stream) - launches on different streams must not share it - grown on
    // demand and kept (a handful of entries: one process drives one GPU with one or two streams)
    struct OvfBuf { int* p = nullptr; long long cap = 0; };
    static std::mutex ovf_mu;
    static std::map<std::pair<int, cudaStream_t>, OvfBuf> ovf_map;
    int* ovf = nullptr;  // [4 ints: counter + padding][cap row ids]
    {
      int dev = 0;
      EDGL_CUDA(cudaGetDevice(&dev));
      std::lock_guard<std::mutex> lock(ovf_mu);
      OvfBuf& ob = ovf_map[std::make_pair(dev, st)];
      if (B > ob.cap) {
        if (ob.p) EDGL_CUDA(cudaFree(ob.p));  // synchronises: the old list is no longer in use
        ob.p = nullptr;
        ob.cap = 0;
        EDGL_CUDA(cudaMalloc(&ob.p, ((size_t)B + 4) * sizeof(int)));
        ob.cap = B;
      }
      ovf = ob.p;
    }
    unsigned int* cnt = reinterpret_cast<unsigned int*>(ovf);
    EDGL_CUDA(cudaMemsetAsync(cnt, 0, sizeof(unsigned int), st));
    const bool vec = (ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(logits) & 15) == 0) && ld >= ((N + 3) & ~3);
    if (vec)
      topk_warp2_kernel<true><<<(unsigned)((B + 7) / 8), 256, 0, st>>>(logits, ld, B, N, K, col_offset, out_stride, idx,
                                                                       val, ovf + 4, cnt);
    else
      topk_warp2_kernel<false><<<(unsigned)((B + 7) / 8), 256, 0, st>>>(logits, ld, B, N, K, col_offset, out_stride, idx,
                                                                        val, ovf + 4, cnt);
    EDGL_LAUNCH_CHECK();
    const int KP2 = next_pow2(K);
    const size_t smem2 = (size_t)(KP2 > kCandCap ? KP2 : kCandCap) * 8;
    topk_list_kernel<<<296, 256, smem2, st>>>(logits, ld, N, K, KP2, col_offset, out_stride, idx, val, ovf + 4, cnt);
    EDGL_LAUNCH_CHECK();
    return 0;
  }
  if (use_warp && !run_if && K <= 128 && N <= 32 * 96 && B >= 64) {
    const unsigned grid = (unsigned)((B + 7) / 8);
    if (N <= 32 * 32) topk_warp_kernel<32><<<grid, 256, 0, st>>>(logits, ld, B, N, K, col_offset, out_stride, idx, val, pp);
    else if (N <= 32 * 72) topk_warp_kernel<72><<<grid, 256, 0, st>>>(logits, ld, B, N, K, col_offset, out_stride, idx, val, pp);
    else topk_warp_kernel<96><<<grid, 256, 0, st>>>(logits, ld, B, N, K, col_offset, out_stride, idx, val, pp);
    EDGL_LAUNCH_CHECK();
    return 0;
  }
  const int KP = next_pow2(K);
  const size_t smem = (size_t)(KP > kCandCap ? KP : kCandCap) * 8;
  topk_kernel<<<B, 256, smem, st>>>(logits, ld, N, K, KP, col_offset, out_stride, idx, val, pp, run_if);
  EDGL_LAUNCH_CHECK();
  return 0;
}

// ---- exact top-K over the candidate list of a row (TopkFilter epilogue of the logits GEMM).  The list holds every
// column whose logit is >= a lower bound of the row's K-th largest UNMASKED logit, so the top-K of the unmasked
// candidates is the top-K of the row.  The ids of the row's own sequence (Base.py:156-163) are dropped through a
// small open-addressing hash in shared memory.  A row whose list overflowed, or that holds fewer than K unmasked
// candidates (ties / constant rows make the bound useless), raises *redo: the caller's predicated launches then
// recompute the whole chunk through the materialised logits.
__global__ void __launch_bounds__(256) topk_select_kernel(const unsigned long long* __restrict__ cand,
                                                          const unsigned int* __restrict__ cnt, int cap, int K,
                                                          int col_offset, const int64_t* __restrict__ seen, int seen_len,
                                                          long long seen_stride, long long col0, long long col1,
                                                          long long out_stride, int32_t* __restrict__ idx_out,
                                                          float* __restrict__ val_out, int* __restrict__ redo,
                                                          int hash_size) {
  extern __shared__ __align__(16) unsigned long long s[];  // [np <= cap] words, then the hash [hash_size] u32
  const int tid = threadIdx.x;
  const long long row = blockIdx.x;
  const unsigned int n = cnt[row];
  if (n > (unsigned int)cap || n < (unsigned int)K) {  // block-uniform
    if (tid == 0) atomicExch(redo, 1);
    return;
  }
  int np = 128;
  while (np < (int)n) np <<= 1;
  uint32_t* hash = reinterpret_cast<uint32_t*>(s + cap);
  for (int i = tid; i < hash_size; i += 256) hash[i] = 0xffffffffu;
  __syncthreads();
  if (seen) {
    for (int l = tid; l < seen_len; l += 256) {
      const long long id = seen[row * seen_stride + l];
      if (id >= col0 && id < col1) {
        const uint32_t v = (uint32_t)(id - col0);
        uint32_t hpos = (v * 2654435761u) & (uint32_t)(hash_size - 1);
        while (true) {
          const uint32_t old = atomicCAS(&hash[hpos], 0xffffffffu, v);
          if (old == 0xffffffffu || old == v) break;
          hpos = (hpos + 1) & (uint32_t)(hash_size - 1);
        }
      }
    }
  }
  __syncthreads();
  const unsigned long long* src = cand + row * cap;
  for (int i = tid; i < np; i += 256) {
    unsigned long long c = 0ull;  // below every real word (the key of -inf is 0x007fffff)
    if (i < (int)n) {
      c = src[i];
      const uint32_t v = 0xffffffffu - (uint32_t)(c & 0xffffffffull);
      uint32_t hpos = (v * 2654435761u) & (uint32_t)(hash_size - 1);
      while (true) {
        const uint32_t hv = hash[hpos];
        if (hv == v) { c = 0ull; break; }
        if (hv == 0xffffffffu) break;
        hpos = (hpos + 1) & (uint32_t)(hash_size - 1);
      }
    }
    s[i] = c;
  }
  bitonic_desc_hybrid(s, np);
  if (s[K - 1] == 0ull) {  // fewer than K unmasked candidates (block-uniform: read after the sort's last barrier)
    if (tid == 0) atomicExch(redo, 1);
    return;
  }
  write_topk(s, K, K, col_offset, idx_out + row * out_stride, val_out + row * out_stride);
}

int launch_topk_select(const unsigned long long* cand, const unsigned int* cnt, int cap, int B, int K, int col_offset,
                       const int64_t* seen, int seen_len, long long seen_stride, long long col0, long long col1,
                       long long out_stride, int32_t* idx, float* val, int* redo, cudaStream_t st) {
  EDGL_REQUIRE(K >= 1 && K <= cap && cap >= 128 && (cap & (cap - 1)) == 0, "topk_select: bad K / cap (%d / %d)", K, cap);
  if (B == 0) return 0;
  if (out_stride == 0) out_stride = K;
  int hs = 256;
  while (hs < 2 * seen_len) hs <<= 1;
  const size_t smem = (size_t)cap * 8 + (size_t)hs * 4;
  EDGL_CUDA(cudaFuncSetAttribute(topk_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  topk_select_kernel<<<B, 256, smem, st>>>(cand, cnt, cap, K, col_offset, seen, seen_len, seen_stride, col0, col1,
                                           out_stride, idx, val, redo, hs);
  EDGL_LAUNCH_CHECK();
  return 0;
}

// merge: shard g's block starts at cand_*[g * shard_stride], its rows are row_stride apart; idx < 0 = padding.
__global__ void __launch_bounds__(256) topk_merge_kernel(const float* __restrict__ cv, const int32_t* __restrict__ ci,
                                                         int G, int Bt, int K, int KP, long long shard_stride,
                                                         long long row_stride, int32_t* __restrict__ idx_out,
                                                         float* __restrict__ val_out) {
  extern __shared__ __align__(16) unsigned long long cand[];
  const int row = blockIdx.x;
  for (int i = threadIdx.x; i < KP; i += 256) {
    unsigned long long c = 0ull;
    if (i < G * K) {
      const int g = i / K, j = i % K;
      const long long o = (long long)g * shard_stride + (long long)row * row_stride + j;
      const int32_t id = ci[o];
      if (id >= 0) c = compose(f2key(cv[o]), (uint32_t)id);
    }
    cand[i] = c;
  }
  bitonic_desc(cand, KP);
  for (int i = threadIdx.x; i < K; i += 256) {
    const unsigned long long c = cand[i];
    const long long o = (long long)row * K + i;
    if (c != 0ull) {
      idx_out[o] = (int32_t)(0xffffffffu - (uint32_t)(c & 0xffffffffull));
      val_out[o] = key2f((uint32_t)(c >> 32));
    } else {
      idx_out[o] = -1;
      val_out[o] = -INFINITY;
    }
  }
}

// Merge of per-shard lists that are already SORTED (value desc, index asc; padding last) - what edgl_logits_topk
// writes: every candidate's final position is its position in its own list plus, for every other list, the number
// of entries that beat it (a binary search), so no sort is needed.  Unsorted input (block-uniform check) takes the
// bitonic path of topk_merge_kernel.  4096 rows x 8 shards: 0.215 ms (sort of 1024 entries per row) -> see DESIGN.md.
__global__ void __launch_bounds__(256) topk_merge_rank_kernel(const float* __restrict__ cv, const int32_t* __restrict__ ci,
                                                              int G, int Bt, int K, int KP, long long shard_stride,
                                                              long long row_stride, int32_t* __restrict__ idx_out,
                                                              float* __restrict__ val_out) {
  extern __shared__ __align__(16) unsigned long long cand[];  // [KP] >= G * K
  const int row = blockIdx.x, n = G * K;
  int bad = 0;
  for (int i = threadIdx.x; i < KP; i += 256) {
    unsigned long long c = 0ull;
    if (i < n) {
      const int g = i / K, j = i - g * K;
      const long long o = (long long)g * shard_stride + (long long)row * row_stride + j;
      const int32_t id = ci[o];
      if (id >= 0) c = compose(f2key(cv[o]), (uint32_t)id);
    }
    cand[i] = c;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += 256) {
    const int j = i % K;
    if (j + 1 < K && cand[i] < cand[i + 1]) bad = 1;
  }
  if (__syncthreads_or(bad)) {  // not sorted: exact general path
    bitonic_desc(cand, KP);
    for (int i = threadIdx.x; i < K; i += 256) {
      const unsigned long long c = cand[i];
      const long long o = (long long)row * K + i;
      idx_out[o] = c != 0ull ? (int32_t)(0xffffffffu - (uint32_t)(c & 0xffffffffull)) : -1;
      val_out[o] = c != 0ull ? key2f((uint32_t)(c >> 32)) : -INFINITY;
    }
    return;
  }
  for (int i = threadIdx.x; i < K; i += 256) {  // padding first; real entries overwrite it after the barrier
    idx_out[(long long)row * K + i] = -1;
    val_out[(long long)row * K + i] = -INFINITY;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += 256) {
    const unsigned long long c = cand[i];
    if (c == 0ull) continue;
    const int g = i / K;
    int rank = i - g * K;
    for (int g2 = 0; g2 < G; ++g2) {
      if (g2 == g) continue;
      const unsigned long long* lst = cand + g2 * K;
      int lo = 0, hi = K;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (lst[mid] > c) lo = mid + 1; else hi = mid;
      }
      rank += lo;
      if (rank >= K) break;
    }
    if (rank < K) {
      idx_out[(long long)row * K + rank] = (int32_t)(0xffffffffu - (uint32_t)(c & 0xffffffffull));
      val_out[(long long)row * K + rank] = key2f((uint32_t)(c >> 32));
    }
  }
}

int launch_topk_merge(const float* cand_val, const int32_t* cand_idx, int G, int Bt, int K, long long shard_stride,
                      long long row_stride, int32_t* idx, float* val, cudaStream_t st) {
  if (row_stride == 0) row_stride = K;
  if (shard_stride == 0) shard_stride = (long long)Bt * row_stride;
  EDGL_REQUIRE(G >= 1 && K >= 1 && (long long)G * K <= 16384, "topk_merge: G*K must be <= 16384");
  if (Bt == 0) return 0;
  const int KP = next_pow2(G * K);
  static const bool old_merge = getenv("EDGL_MERGE_SORT") != nullptr;
  auto kern = old_merge ? topk_merge_kernel : topk_merge_rank_kernel;
  const size_t smem = (size_t)KP * 8;
  if (smem > 48 * 1024) EDGL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<Bt, 256, smem, st>>>(cand_val, cand_idx, G, Bt, K, KP, shard_stride, row_stride, idx, val);
  EDGL_LAUNCH_CHECK();
  return 0;
}

// ---- fused exchange 1: packed [y | seqs_i] rows of this rank -> slot `rank` of EVERY peer's gathered buffer
__global__ void __launch_bounds__(256) put_rows_kernel(const float* __restrict__ y, long long ldy,
                                                       const int64_t* __restrict__ ids, int L, int d, int B,
                                                       const long long* __restrict__ peer_rows, int G, int rank,
                                                       const long long* __restrict__ peer_flags, uint32_t epoch,
                                                       unsigned int* counter) {
  const int b = blockIdx.x;
  const int W = d + 2 * L;
  const float* idf = reinterpret_cast<const float*>(ids + (long long)b * L);  // raw bytes of the int64 ids
  for (int t = threadIdx.x; t < W; t += blockDim.x) {
    const float v = t < d ? y[b * ldy + t] : idf[t - d];
    const long long off = ((long long)rank * B + b) * W + t;
    for (int g = 0; g < G; ++g) reinterpret_cast<float*>(peer_rows[g])[off] = v;
  }
  p2p_signal_when_grid_done(counter, peer_flags, G, rank, epoch);
}

int launch_put_rows(const float* y, long long ldy, const int64_t* ids, int L, int d, int B, const long long* peer_rows,
                    int G, int rank, const long long* peer_flags, uint32_t epoch, unsigned int* counter,
                    cudaStream_t st) {
  if (B == 0) return 0;
  put_rows_kernel<<<B, 256, 0, st>>>(y, ldy, ids, L, d, B, peer_rows, G, rank, peer_flags, epoch, counter);
  EDGL_LAUNCH_CHECK();
  return 0;
}

// spin until every one of the G flags has reached `epoch` (wrap-around safe)
__global__ void wait_flags_kernel(const uint32_t* flags, int G, uint32_t epoch) {
  const int g = threadIdx.x;
  if (g < G)
    while ((int)(ld_acquire_sys(flags + g) - epoch) < 0) __nanosleep(64);
}

int launch_wait_flags(const uint32_t* flags, int G, uint32_t epoch, cudaStream_t st) {
  wait_flags_kernel<<<1, 32, 0, st>>>(flags, G, epoch);
  EDGL_LAUNCH_CHECK();
  return 0;
}

}  // namespace edgl
