// common.cuh - shared helpers for libeasydgl_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <string>

namespace edgl {

// thread-local last error message (edgl_last_error)
std::string& last_error();
int set_error(int code, const char* fmt, ...);

extern std::atomic<long long> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define EDGL_CUDA(expr)                                                                          \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess)                                                                       \
      return ::edgl::set_error(-3, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                               __LINE__);                                                        \
  } while (0)

#define EDGL_LAUNCH_CHECK()                                                                      \
  do {                                                                                           \
    ::edgl::count_launch();                                                                      \
    cudaError_t _e = cudaGetLastError();                                                         \
    if (_e != cudaSuccess)                                                                       \
      return ::edgl::set_error(-3, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e),    \
                               __FILE__, __LINE__);                                              \
  } while (0)

#define EDGL_REQUIRE(cond, ...)                                  \
  do {                                                           \
    if (!(cond)) return ::edgl::set_error(-1, __VA_ARGS__);      \
  } while (0)

#define EDGL_TRY(expr)          \
  do {                          \
    int _r = (expr);            \
    if (_r != 0) return _r;     \
  } while (0)

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

#ifdef __CUDACC__
// Running maximum of |x| over a tensor, kept as the bit pattern of a non-negative float in a device word (zeroed at
// the start of a forward).  Producers call it once per warp with the warp's maximum; the racy read skips the atomic
// once the slot is already at least as large (same-address atomics would otherwise serialise the whole grid).
// Consumer: the scaled 3xFP16 GEMM (gemm_f16.cu) derives its activation scale from it.
__device__ __forceinline__ void amax_publish(unsigned int* slot, float warp_local_max, int lane) {
  const unsigned int bits = __reduce_max_sync(0xffffffffu, __float_as_uint(fabsf(warp_local_max)));
  if (lane == 0 && bits > *reinterpret_cast<volatile unsigned int*>(slot)) atomicMax(slot, bits);
}
#endif

// ------------------------------------------------------------------ kernel launchers (one per .cu)
enum Act { ACT_NONE = 0, ACT_GELU = 1, ACT_RELU = 2 };

// LayerNorm folded into the tensor-core dense layers (Base.layernorm over (L, C) jointly, Base.py:12-67): the producer
// of a tensor emits per-row partial sums in its epilogue, launch_ln_finalize turns them into (mean, rstd) per
// sequence, and the consumers normalise on the fly - no LayerNorm pass reads or writes the activations:
//  * as the A operand: the splitter warps, which touch every A element in shared memory anyway, write
//    (x - mean) rstd gamma + beta in place before they split it;
//  * as the residual: the epilogue applies the same formula to the residual rows it loads.
struct LnEpi {
  const float2* a_rs = nullptr;  // per-sequence (mean, rstd) of the LayerNorm applied to A
  const float* a_g = nullptr;    // [K] gamma
  const float* a_b = nullptr;    // [K] beta
  const float2* r_rs = nullptr;  // per-sequence (mean, rstd) of the LayerNorm applied to the residual rows
  const float* r_g = nullptr;    // [N] gamma
  const float* r_b = nullptr;    // [N] beta
  float2* stats = nullptr;       // [M][2 * ceil(N / BN)] row partial (sum, sum of squares) of the stored values
  int L = 1;                     // rows per sequence
  int last_only = 0;             // store only the rows with row % L == L-1, at C[(row / L) * ldc]
  bool any() const { return a_rs || r_rs || stats || last_only; }
};
int ln_stats_parts(int N);  // number of partials per row the tensor-core GEMM writes for N columns

// ---- precision ablation (tools/ablation.py).  Every contraction runs as three tensor-core products
// (A_lo B_hi + A_hi B_lo + A_hi B_hi); these switches drop individual products so that their effect on the logits and
// on the top-K sets can be measured.  Off unless the environment variables are set (read per launch; a skipped MMA is
// a predicate in the single MMA-issuing thread of the GEMMs and a separate template instantiation of the attention
// kernel: the default path is unchanged).
//   EDGL_ABL_GEMM = six digits 0..3 for the dense layers qkvt, ao, ff1, ff2, tr, logits: bit 0 drops A_lo W_hi,
//                   bit 1 drops A_hi W_lo
//   EDGL_ABL_ATTN = bit mask: 1 Q_lo K_hi, 2 Q_hi K_lo, 4 P_lo [T|V]_hi, 8 P_hi [T|V]_lo, 16 H_lo W1_hi, 32 H_hi W1_lo,
//                   64 lam_lo M  (attention_f16_kernel, L <= 104 instantiation only)
extern int g_edgl_stage;      // pipeline stage of the launch in flight (api.cu mark())
int ablation_gemm_bits();     // bits of EDGL_ABL_GEMM for g_edgl_stage, 0 when unset / not a dense stage
int ablation_attn_mask();     // EDGL_ABL_ATTN, 0 when unset

// order-preserving unsigned keys of fp32 values and the 64-bit (key, ~index) words the ranking kernels sort:
// descending order of the word = (value descending, index ascending), tf.nn.top_k's order (Base.py:181)
__device__ __forceinline__ uint32_t f2key(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}
__device__ __forceinline__ unsigned long long compose(uint32_t key, uint32_t idx) {
  return ((unsigned long long)key << 32) | (unsigned long long)(0xffffffffu - idx);
}

// Top-K candidate filter in the epilogue of the logits GEMM (gemm_tc.cu; api.cu logits_topk): instead of storing
// C, every value >= thr[row] - a lower bound of the row's K-th largest unmasked logit, taken from a column sample -
// is appended to the row's candidate list as a (key, column) word.  cnt[row] counts every hit, also those beyond
// `cap` (the caller detects the overflow and falls back to the materialised path).
struct TopkFilter {
  const float* thr = nullptr; long long thr_stride = 0;
  unsigned long long* cand = nullptr; int cap = 0;  // [M][cap]
  unsigned int* cnt = nullptr;                      // [M], zeroed by the caller
};
// (mean, rstd)[b] from the row partials of B sequences of L rows x C columns; optionally y[b] = LN(x_last[b]) with
// x_last [B, C] (the stored last rows), gamma, beta
int launch_ln_finalize(const float2* parts, int nparts, int B, int L, int C, float2* rs, const float* x_last,
                       const float* gamma, const float* beta, float* y, unsigned int* y_amax, cudaStream_t st,
                       long long y_stride = 0);
// C[M,N] (ldc) = act(A[M,K] (lda) @ W + bias[N] + pbias[(m % pperiod), N]) + R[M,N] (ldr)
// W is [K,N] (ldw) row-major, or, if w_is_nk, [N,K] (ldw) row-major (C = A @ W^T).
// col0_const: if non-null-flag set, global column (col_offset+0)==0 gets exactly bias (acc ignored)
struct GemmArgs {
  const float* A = nullptr; int lda = 0;
  const float* W = nullptr; int ldw = 0; bool w_is_nk = false;
  const float* Wlo = nullptr;  // optional, w_is_nk only: W - tf32(W) in the layout of W (constant weights, made at commit)
  // scaled 3xFP16 path (gemm_f16.cu), all three needed: the weight's fp16 hi/lo/scale buffer (launch_w_split_f16),
  // the running max|A| published by A's producer, and optionally where to publish max|C| for the next layer
  const void* W16 = nullptr;
  const unsigned int* a_amax = nullptr;
  unsigned int* c_amax = nullptr;
  float* C = nullptr; int ldc = 0;
  int M = 0, N = 0, K = 0;
  const float* bias = nullptr;                 // [N] or null
  const float* pbias = nullptr; int pperiod = 0;  // [pperiod, N] or null
  const float* R = nullptr; int ldr = 0;       // residual or null
  int act = ACT_NONE;
  bool zero_wrow0 = false;  // w_is_nk only: treat W row 0 (= output column 0) as all-zero (zero_pad table)
  LnEpi ln;                 // tensor-core path only (3xTF32 kernel): fused LayerNorm pieces, see above
  TopkFilter flt;           // tensor-core path only (3xTF32 kernel): candidate filter instead of the store of C
  const int* run_if = nullptr;  // device flag or null: the kernel does nothing when *run_if == 0 (fallback launches)
};
int launch_gemm(const GemmArgs& a, cudaStream_t st);
bool gemm_tc_supported(const GemmArgs& a);
int launch_gemm_tc(const GemmArgs& a, cudaStream_t st);
int launch_transpose(const float* in, int rows, int cols, float* out, cudaStream_t st);
int launch_tf32_lo(const float* in, long long n, float* out, cudaStream_t st);
bool gemm_f16_supported(const GemmArgs& a);
int launch_gemm_f16(const GemmArgs& a, cudaStream_t st);
size_t w16_bytes(long long n);  // buffer size for the fp16 hi / lo / scale copies of an n-element weight
int launch_w_split_f16(const float* w, long long n, void* buf, cudaStream_t st);
int launch_absmax(const float* x, long long n, unsigned int* slot, cudaStream_t st);  // out = in - (in with 13 low mantissa bits cleared)

struct EmbedArgs {
  int model;  // 0 EasyDGL, 1 CTSMA
  const int64_t* ids; const float* ts; int B, L, ts_len, d, E;
  float time_scale; int64_t mask_id;
  const float* item_table; int num_rows;       // raw variable, row 0 treated as zero
  const float* pos_table;                      // [L,d]
  const float* mark_embs;                      // [E,d] raw (row 0 treated as zero); EasyDGL only
  const uint8_t* mark_table8; int mark_rows;   // [mark_rows,E]
  const float* tscale;                         // [d/2] sinusoid scales (fp32)
  // outputs (any may be null)
  float* X0; int ldx0;                         // full concat [B*L, 3d | 2d]
  float* Xa; int ldxa;                         // fused input: [x+tcode | mark-count histogram] (EasyDGL), width d+E
  unsigned int* xa_amax;                       // optional: running max |Xa| (amax_publish), or null
  float* spans;                                // [B*L]
  uint8_t* marks;                              // [B*L,E]
  uint8_t* kmask;                              // [B*L]
};
int launch_embed(const EmbedArgs& a, cudaStream_t st);
int launch_time_code(const float* ts, const float* tscale, long long rows, int d, float* out, cudaStream_t st);
int launch_time_function_code(const float* x, const float* freq, const float* phase, long long n, int d, float* out,
                              cudaStream_t st);
int launch_lookup(const float* table, int vocab, int d, int zero_pad, float scale, const int64_t* ids,
                  long long n, float* out, cudaStream_t st);
int launch_mark_table_to_u8(const int64_t* src, long long n, uint8_t* dst, int* err_flag, int E, cudaStream_t st);

struct AttnArgs {
  const float* Q; int ldq;   // [B*L, >=d] Q columns start at Q
  const float* K; int ldk;
  const float* V; int ldv;
  const float* T; int ldt;
  const uint8_t* kmask;      // [B*L]
  const float* spans;        // [B*L]
  const uint8_t* marks;      // [B*L,E]
  const float* R; int ldr;   // residual (first d columns) or null
  const float* int_w; const float* int_b; const float* int_weight; const float* int_scaling;
  float* O; int ldo;         // [B*L, d]
  float* lam;                // [h*B, L, E] or null
  int B, L, d, h, E;
  bool causal, diag_one;
  const void* mlp_pack = nullptr;  // attn_f16.cu: constants of the intensity MLP packed at commit, or null
  unsigned int* out_amax = nullptr;  // attn_f16.cu / attn_tc2.cu: running max |O| (amax_publish), or null
  bool* amax_published = nullptr;    // host flag, set by launch_attention: did the kernel that ran publish out_amax?
  const void* mlp_pack2 = nullptr;   // attn_tc2.cu: constants of the intensity MLP (tcgen05 operand layout), or null
  float* dbg = nullptr;              // attn_tc2.cu, debugging only: [items][8 phases][128 rows][256 cols] intermediates
  long long* prof = nullptr;         // attn_tc2.cu, profiling only: [grid][2 slots][16] per-phase cycle counters
};
int launch_attention(const AttnArgs& a, cudaStream_t st);
// scaled 3xFP16 mma.sync kernel (attn_f16.cu): 0 = launched, 1 = shape not covered, <0 = error
int launch_attention_f16(const AttnArgs& a, cudaStream_t st);
// key-streaming two-pass variant (attn_f16_long.cu): dh in {16, 32}, E = 16, any L
int launch_attention_f16_long(const AttnArgs& a, cudaStream_t st);
size_t attention_f16_pack_bytes(int dh, int E);  // 0 if the shape is not covered
int launch_attention_f16_pack(const float* int_w, const float* int_b, const float* int_weight, const float* int_scaling,
                              int dh, int E, void* pack, cudaStream_t st);
int launch_attention_tc(const AttnArgs& a, cudaStream_t st);
// tcgen05 / TMA kernel with two items in flight and column-split rows (attn_tc2.cu): 0 = launched, 1 = not covered
int launch_attention_tc2(const AttnArgs& a, cudaStream_t st);
size_t attention_tc2_pack_bytes(int dh, int E);  // 0 if the shape is not covered
int launch_attention_tc2_pack(const float* int_w, const float* int_b, const float* int_weight, const float* int_scaling,
                              int dh, int E, void* pack, cudaStream_t st);
// launch_attention with an explicit kernel choice: 'd' default, '2' tc2, 'f' f16 mma.sync, 't' tc (v1), 'm' mma tf32,
// 's' CUDA cores
int launch_attention_mode(const AttnArgs& a, cudaStream_t st, char mode);
int launch_intensity(const float* H, const float* spans, const uint8_t* marks, const float* int_w,
                     const float* int_b, const float* int_weight, const float* int_scaling, int B, int L,
                     int h, int dh, int E, float* G, float* lam, cudaStream_t st);

// LayerNorm over (L,C) jointly per sample. last_only: write only row L-1 to out [B,C].
int launch_layernorm(const float* x, const float* gamma, const float* beta, int B, int L, int C, float* out,
                     bool last_only, cudaStream_t st, unsigned int* out_amax = nullptr, long long out_stride = 0);

// training-mode forward (train.cu)
int launch_gather_rows(const float* Y, const int64_t* pos, int L, int M, int d, long long rows, float* out, int* err,
                       cudaStream_t st);
int launch_ce_rows(const float* logits, int ld, int N, const int64_t* labels, long long row0, int rows, float* pe,
                   float* wt, int* err, cudaStream_t st);
int launch_tpp_terms(const float* lam, const int64_t* positions, const int64_t* labels, const uint8_t* mark8,
                     int mark_rows, const float* ts, int ts_len, int B, int L, int M, int heads, int E, float* ell,
                     float* nu, float* cnt, cudaStream_t st);
int launch_reduce(const float* x, long long n, int squares, double scale, double* out, int accumulate, cudaStream_t st);
int launch_loss_combine(const double* acc, int num_blocks, double ct_scale, float* loss_out, cudaStream_t st);

// time_attn.cu: attention cores of the Ti / Tf / Tg baseline layers (temporal.py:15-264) and their helpers
int launch_time_attention(const float* Q, const float* K, const float* V, const uint8_t* kmask, const uint8_t* qmask,
                          const float* pos_k, const float* pos_v, int mode, const void* intervals, const float* tk,
                          const float* tv, int vocab, const float* freq, const float* phase, const float* U, float* TC,
                          const float* R, int B, int Tq, int Tk, int C, int h, int causal, float* out, cudaStream_t st);
int launch_row_nonzero(const float* x, long long rows, int C, uint8_t* out, cudaStream_t st);
int launch_rownorm(const float* x, const float* gamma, const float* beta, long long rows, int C, float eps, float* out,
                   cudaStream_t st);

int launch_mask_seen(float* logits, int ld, int B, const int64_t* ids, int seen_len, long long seen_stride,
                     long long col0, long long col1, cudaStream_t st, const int* run_if = nullptr);
// Fused candidate exchange over peer memory (multi-GPU): when `dest` is set the top-K kernel writes row R's
// candidates into the receive buffer of rank R / rows_per_dest and, after the last launch of a step, raises
// `epoch` in every peer's flag array.
struct TopkP2P {
  const long long* dest;        // device array [G]: base address of each rank's candidate buffer [G][B][2][K] int32
  const long long* peer_flags;  // device array [G]: address of each rank's uint32 flag array [G]
  unsigned int* counter;        // device scratch (grid-done counter), zero-initialised
  int rows_per_dest, my_rank, row_base, G, signal;
  uint32_t epoch;
};
int launch_topk(const float* logits, int ld, int B, int N, int K, int col_offset, long long out_stride, int32_t* idx,
                float* val, cudaStream_t st, const TopkP2P* p2p = nullptr, const int* run_if = nullptr);
// exact top-K of the candidate lists a TopkFilter epilogue produced (ids of the row's own sequence dropped: seen-mask,
// Base.py:156-163); rows whose list overflowed `cap` or holds fewer than K unmasked entries raise *redo
int launch_topk_select(const unsigned long long* cand, const unsigned int* cnt, int cap, int B, int K, int col_offset,
                       const int64_t* seen, int seen_len, long long seen_stride, long long col0, long long col1,
                       long long out_stride, int32_t* idx, float* val, int* redo, cudaStream_t st);
int launch_put_rows(const float* y, long long ldy, const int64_t* ids, int L, int d, int B, const long long* peer_rows,
                    int G, int rank, const long long* peer_flags, uint32_t epoch, unsigned int* counter,
                    cudaStream_t st);
int launch_wait_flags(const uint32_t* flags, int G, uint32_t epoch, cudaStream_t st);
int launch_topk_merge(const float* cand_val, const int32_t* cand_idx, int G, int Bt, int K, long long shard_stride,
                      long long row_stride, int32_t* idx, float* val, cudaStream_t st);

}  // namespace edgl
