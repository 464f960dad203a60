// api.cu - the C ABI (include/easydgl_b200.h): handle, weight binding, workspace, and the kernel
// pipelines for EasyDGL.__call__ (EasyDGL.py:69-151), CTSMA.__call__ (CTSMA.py:46-91) and the ranking
// part of Sequential.eval (Base.py:150-181).
#include <math.h>
#include <stdarg.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/easydgl_b200.h"
#include "common.cuh"

namespace edgl {

std::atomic<long long> g_launches{0};

std::string& last_error() {
  static thread_local std::string e;
  return e;
}

int set_error(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  last_error() = buf;
  return code;
}

// activation-maximum slots (edgl_handle::amax)
enum { AMAX_XA = 0, AMAX_ATT = 1, AMAX_LN1 = 2, AMAX_FF1 = 3, AMAX_LN2 = 4, AMAX_Y = 5, AMAX_COUNT = 16 };

struct Tensor {
  const void* p = nullptr;
  long long numel = 0;
};

// pipeline stages (one kernel launch each) for the optional CUDA-event profile (edgl_profile*)
enum Stage {
  ST_EMBED = 0, ST_LN_IN, ST_QKVT_GEMM, ST_ATTENTION, ST_AO_GEMM, ST_LN_ATT, ST_FF1_GEMM, ST_FF2_GEMM, ST_LN_FF,
  ST_TR_GEMM, ST_LN_OUT, ST_LOGITS_GEMM, ST_MASK_SEEN, ST_TOPK, ST_END, ST_COUNT
};
const char* const kStageNames[ST_COUNT] = {"embed", "ln_in", "qkvt_gemm", "attention", "ao_gemm", "ln_att",
                                           "ff1_gemm", "ff2_gemm", "ln_ff", "tr_gemm", "ln_out", "logits_gemm",
                                           "mask_seen", "topk", "end"};

int g_edgl_stage = -1;
int ablation_gemm_bits() {
  const char* e = getenv("EDGL_ABL_GEMM");
  if (!e) return 0;
  int idx = -1;
  switch (g_edgl_stage) {
    case ST_QKVT_GEMM: idx = 0; break;
    case ST_AO_GEMM: idx = 1; break;
    case ST_FF1_GEMM: idx = 2; break;
    case ST_FF2_GEMM: idx = 3; break;
    case ST_TR_GEMM: idx = 4; break;
    case ST_LOGITS_GEMM: idx = 5; break;
    default: return 0;
  }
  if ((int)strlen(e) <= idx || e[idx] < '0' || e[idx] > '3') return 0;
  return e[idx] - '0';
}
int ablation_attn_mask() {
  const char* e = getenv("EDGL_ABL_ATTN");
  return e ? atoi(e) : 0;
}

}  // namespace edgl

using namespace edgl;

struct edgl_handle {
  edgl_config cfg;
  int dev = 0;
  int L = 0, d = 0, h = 0, dh = 0, E = 0, N1 = 0, K = 0, ts_len = 0;
  int Ka = 0;            // width of the fused block-0 input: d + E (EasyDGL) or 2d (CTSMA)
  long long c0 = 0, c1 = 0;  // logit columns owned by this handle
  std::map<std::string, Tensor> mt;                // model-level tensors
  std::vector<std::map<std::string, Tensor>> bt;   // per-block tensors
  bool committed = false;
  // derived (owned)
  float* tscale = nullptr;
  uint8_t* mark8 = nullptr;
  int* flag = nullptr;
  int* topk_redo = nullptr;  // device flag of the fused logits + top-K path: a row needs the materialised fallback
  unsigned int* p2p_counter = nullptr;  // grid-done counters of the fused-exchange kernels (zero-initialised)
  float* bias_full = nullptr;  // [N1] = concat([-1000], output_bias)  (Base.py:110)
  float* table_lo = nullptr;   // [c1-c0, d] tf32 lo part of the owned item-table rows (logits GEMM), or null
  float* wfold0 = nullptr;     // EasyDGL block 0: [Ka,4d]
  float* pbias0 = nullptr;     // EasyDGL block 0: [L,4d] = pos_embs @ W[d:2d] + b
  std::vector<float*> wkvt, bkvt;  // CTSMA: packed [Cin,3d], [3d]
  std::vector<unsigned char*> mlp_pack;  // per block: intensity-MLP constants for attn_f16.cu (null if not covered)
  std::vector<unsigned char*> mlp_pack2;  // per block: the same constants in the tcgen05 operand layout of attn_tc2.cu
  // scaled 3xFP16 dense layers (gemm_f16.cu): fp16 hi/lo/scale copies of the K-major kernels, keyed like btT / mtT,
  // and the running activation maxima the producers publish (zeroed at the start of every encode)
  bool f16_gemm = false;
  int f16_mask = 1;  // which dense layers take it: bit 0 QKVT, 1 attention-out, 2 FF1, 3 FF2, 4 transform, 5 logits
  std::vector<std::map<std::string, unsigned char*>> bt16;
  std::map<std::string, unsigned char*> mt16;  // "tr_w", "wfold0", "table"
  unsigned int* amax = nullptr;                // [16]: AMAX_* slots
  // K-major ([N,K]) copies of every dense kernel for the tensor-core GEMM (made in edgl_commit)
  float* wfold0T = nullptr;                               // [4d, Ka]
  std::vector<std::map<std::string, float*>> btT;         // per block: name -> [N,K]
  std::map<std::string, float*> mtT;                      // model level (tr_w)
  // LayerNorm folded into the tensor-core dense layers (LnEpi, common.cuh): EasyDGL, d % 16 == 0
  bool ln_fuse = false;
  float2* ln_parts = nullptr;                      // [rows][parts] row partial sums
  float2 *ln_rs1 = nullptr, *ln_rs2 = nullptr, *ln_rs3 = nullptr;  // [max_batch] (mean, rstd)
  float* tr_last = nullptr;                        // [max_batch, d] last rows of the transform layer
  // training-mode forward (train.cu): allocated on first use (grow-only)
  std::vector<float*> train_lam;   // per block: lam [h * max_batch, L, E]
  float* train_y = nullptr;        // gathered hidden rows [rows_cap, d]
  float *train_pe = nullptr, *train_wt = nullptr, *train_tpp = nullptr;  // [rows_cap], [rows_cap], [3][h * rows_cap]
  long long train_rows_cap = 0;
  double* train_acc = nullptr;     // reduction slots
  // workspace (owned)
  float *xa = nullptr, *p0 = nullptr, *p1 = nullptr, *p2 = nullptr, *qkvt = nullptr, *spans = nullptr, *y = nullptr,
        *logits_ws = nullptr;
  uint8_t *marks = nullptr, *kmask = nullptr;
  long long ws_rows = 0;  // rows of logits_ws
  // staging for the *_host entry points: two slots, so the upload of batch i+1 and the download of batch i-1 overlap
  // the kernels of batch i (copies on two side streams, ordered by events)
  int64_t* st_ids[2] = {nullptr, nullptr};
  float* st_ts[2] = {nullptr, nullptr};
  int32_t* st_idx[2] = {nullptr, nullptr};
  float* st_val[2] = {nullptr, nullptr};
  cudaStream_t cs_in = nullptr, cs_out = nullptr;
  cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr}, ev_d2h[2] = {nullptr, nullptr};
  bool slot_busy[2] = {false, false}, slot_used[2] = {false, false};
  int next_slot = 0;
  std::vector<void*> owned;
  // optional per-stage CUDA-event profile
  bool prof_on = false;
  std::vector<cudaEvent_t> prof_ev;
  std::vector<int> prof_stage;
  size_t prof_n = 0;
};

namespace {

// record an event BEFORE the stage's kernel is enqueued; the time to the next mark belongs to `stage`
inline void mark(edgl_handle* h, int stage, cudaStream_t st) {
  g_edgl_stage = stage;
  if (!h->prof_on || h->prof_n >= h->prof_ev.size()) return;
  cudaEventRecord(h->prof_ev[h->prof_n], st);
  h->prof_stage[h->prof_n] = stage;
  ++h->prof_n;
}

template <typename T>
int dev_alloc(edgl_handle* h, T** p, size_t n) {
  void* q = nullptr;
  if (n == 0) n = 1;
  cudaError_t e = cudaMalloc(&q, n * sizeof(T));
  if (e != cudaSuccess) {
    cudaGetLastError();
    return set_error(EDGL_ENOMEM, "cudaMalloc of %zu bytes failed: %s", n * sizeof(T), cudaGetErrorString(e));
  }
  h->owned.push_back(q);
  *p = reinterpret_cast<T*>(q);
  return 0;
}

int cin_of(const edgl_handle* h, int block) {
  if (h->cfg.model == EDGL_MODEL_EASYDGL) return block == 0 ? 3 * h->d : h->d;
  return block == 0 ? 2 * h->d : h->d;
}

// expected element count of a named tensor; -1 = unknown name
long long expected_numel(const edgl_handle* h, const std::string& n, int block) {
  const long long d = h->d, E = h->E, dh = h->dh, L = h->L, N1 = h->N1;
  const bool easy = h->cfg.model == EDGL_MODEL_EASYDGL;
  if (block < 0) {
    if (n == "item_embs") return N1 * d;
    if (n == "pos_embs") return L * d;
    if (n == "output_bias") return N1 - 1;
    if (n == "mark_table") return (long long)h->cfg.mark_rows * E;
    if (easy) {
      if (n == "mark_embs") return E * d;
      if (n == "tr_w") return d * d;
      if (n == "tr_b" || n == "tr_ln_g" || n == "tr_ln_b") return d;
    } else {
      if (n == "out_ln_g" || n == "out_ln_b") return d;
    }
    return -1;
  }
  const long long cin = cin_of(h, block);
  if (n == "int_w") return (dh + 1) * dh * E;
  if (n == "int_b") return dh * E;
  if (n == "int_weight") return E * dh;
  if (n == "int_scaling") return E;
  if (easy) {
    if (n == "qkvt_w") return cin * 4 * d;
    if (n == "qkvt_b") return 4 * d;
    if (n == "ao_w") return d * d;
    if (n == "ao_b" || n == "ao_ln_g" || n == "ao_ln_b" || n == "ff2_b" || n == "ff_ln_g" || n == "ff_ln_b") return d;
    if (n == "ff1_w" || n == "ff2_w") return 2 * d * d;
    if (n == "ff1_b") return 2 * d;
  } else {
    if (n == "ln1_g" || n == "ln1_b") return cin;
    if (n == "q_w" || n == "k_w" || n == "v_w" || n == "t_w") return cin * d;
    if (n == "q_b" || n == "k_b" || n == "v_b" || n == "t_b" || n == "ln2_g" || n == "ln2_b" || n == "ff1_b" ||
        n == "ff2_b")
      return d;
    if (n == "ff1_w" || n == "ff2_w") return d * d;
  }
  return -1;
}

const char* const kModelNamesEasy[] = {"item_embs", "pos_embs", "output_bias", "mark_table", "mark_embs",
                                       "tr_w", "tr_b", "tr_ln_g", "tr_ln_b"};
const char* const kModelNamesCtsma[] = {"item_embs", "pos_embs", "output_bias", "mark_table", "out_ln_g", "out_ln_b"};
const char* const kBlockNamesEasy[] = {"int_w", "int_b", "int_weight", "int_scaling", "qkvt_w", "qkvt_b", "ao_w",
                                       "ao_b", "ao_ln_g", "ao_ln_b", "ff1_w", "ff1_b", "ff2_w", "ff2_b",
                                       "ff_ln_g", "ff_ln_b"};
const char* const kBlockNamesCtsma[] = {"int_w", "int_b", "int_weight", "int_scaling", "ln1_g", "ln1_b", "q_w",
                                        "q_b", "k_w", "k_b", "v_w", "v_b", "t_w", "t_b", "ln2_g", "ln2_b",
                                        "ff1_w", "ff1_b", "ff2_w", "ff2_b"};

inline const float* F(const std::map<std::string, Tensor>& m, const char* n) {
  return reinterpret_cast<const float*>(m.at(n).p);
}

int check_ready(const edgl_handle* h, int B, bool need_batch_fit = true) {
  if (!h) return set_error(EDGL_EINVAL, "null handle");
  if (!h->committed) return set_error(EDGL_ESTATE, "edgl_commit has not been called after the last edgl_set_tensor");
  if (B < 0 || (need_batch_fit && B > h->cfg.max_batch))
    return set_error(EDGL_EINVAL, "batch %d outside [0, max_batch=%d]", B, h->cfg.max_batch);
  return 0;
}

int dense(const float* A, int lda, const float* W, int ldw, const float* bias, float* C, int ldc, long long M, int N,
          int K, int act, const float* R, int ldr, cudaStream_t st) {
  GemmArgs g;
  g.A = A; g.lda = lda; g.W = W; g.ldw = ldw; g.C = C; g.ldc = ldc;
  g.M = (int)M; g.N = N; g.K = K; g.bias = bias; g.act = act; g.R = R; g.ldr = ldr;
  return launch_gemm(g, st);
}

// dense layer with a K-major ([N,K]) kernel -> tensor-core path
int dense_nk(const float* A, int lda, const float* Wt, int K, const float* bias, float* C, int ldc, long long M, int N,
             int act, const float* R, int ldr, cudaStream_t st, bool has_lo = true, const void* w16 = nullptr,
             const unsigned int* a_amax = nullptr, unsigned int* c_amax = nullptr, const LnEpi* ln = nullptr) {
  GemmArgs g;
  if (ln) g.ln = *ln;
  g.W16 = w16; g.a_amax = w16 ? a_amax : nullptr; g.c_amax = c_amax;
  g.A = A; g.lda = lda; g.W = Wt; g.ldw = K; g.w_is_nk = true; g.C = C; g.ldc = ldc;
  if (has_lo) g.Wlo = Wt + (size_t)N * K;  // every K-major copy made by edgl_commit is followed by its tf32 lo part
  g.M = (int)M; g.N = N; g.K = K; g.bias = bias; g.act = act; g.R = R; g.ldr = ldr;
  return launch_gemm(g, st);
}

AttnArgs attn_args(const edgl_handle* h, const std::map<std::string, Tensor>& w, const float* qkvt, const uint8_t* kmask,
                   const float* spans, const uint8_t* marks, const float* R, int ldr, float* O, float* lam, int B,
                   bool causal, bool diag_one) {
  AttnArgs a;
  const int d = h->d;
  a.Q = qkvt; a.K = qkvt + d; a.V = qkvt + 2 * d; a.T = qkvt + 3 * d;
  a.ldq = a.ldk = a.ldv = a.ldt = 4 * d;
  a.kmask = kmask; a.spans = spans; a.marks = marks; a.R = R; a.ldr = ldr;
  a.int_w = F(w, "int_w"); a.int_b = F(w, "int_b"); a.int_weight = F(w, "int_weight");
  a.int_scaling = F(w, "int_scaling");
  a.O = O; a.ldo = d; a.lam = lam; a.B = B; a.L = h->L; a.d = d; a.h = h->h; a.E = h->E;
  a.causal = causal; a.diag_one = diag_one;
  a.mlp_pack = nullptr;
  for (size_t i = 0; i < h->bt.size() && i < h->mlp_pack.size(); ++i)
    if (&h->bt[i] == &w) a.mlp_pack = h->mlp_pack[i];
  for (size_t i = 0; i < h->bt.size() && i < h->mlp_pack2.size(); ++i)
    if (&h->bt[i] == &w) a.mlp_pack2 = h->mlp_pack2[i];
  return a;
}

EmbedArgs embed_args(const edgl_handle* h, const int64_t* ids, const float* ts, int B) {
  EmbedArgs e;
  memset(&e, 0, sizeof(e));
  e.model = h->cfg.model; e.ids = ids; e.ts = ts; e.B = B; e.L = h->L; e.ts_len = h->ts_len; e.d = h->d; e.E = h->E;
  e.time_scale = h->cfg.time_scale; e.mask_id = h->cfg.mask_id;
  e.item_table = F(h->mt, "item_embs"); e.num_rows = h->N1; e.pos_table = F(h->mt, "pos_embs");
  e.mark_embs = h->cfg.model == EDGL_MODEL_EASYDGL ? F(h->mt, "mark_embs") : nullptr;
  e.mark_table8 = h->mark8; e.mark_rows = h->cfg.mark_rows; e.tscale = h->tscale;
  return e;
}

// training-mode forward: the encoder also hands out the full LayerNorm'ed hidden states and every block's intensities
struct TrainOut {
  float* Y = nullptr;               // [B*L, d]
  std::vector<float*> lam;          // per block [h*B, L, E], or empty
};

// EasyDGL.__call__ up to y = hidden[:, -1]  (EasyDGL.py:69-146)
int encode_easydgl(edgl_handle* h, const int64_t* ids, const float* ts, int B, float* y, long long ldy, cudaStream_t st,
                   const TrainOut* tr = nullptr) {
  const int d = h->d, L = h->L;
  const long long rows = (long long)B * L;
  EmbedArgs e = embed_args(h, ids, ts, B);
  e.Xa = h->xa; e.ldxa = h->Ka; e.spans = h->spans; e.marks = h->marks; e.kmask = h->kmask;
  const bool f16 = h->f16_gemm;
  unsigned int* const am = h->amax;
  auto w16 = [&](int blk, const char* name, int bit) -> const void* {
    return (f16 && ((h->f16_mask >> bit) & 1)) ? h->bt16[blk].at(name) : nullptr;
  };
  if (f16) {
    EDGL_CUDA(cudaMemsetAsync(am, 0, AMAX_COUNT * sizeof(unsigned int), st));
    e.xa_amax = am + AMAX_XA;
  }
  mark(h, ST_EMBED, st);
  EDGL_TRY(launch_embed(e, st));
  const float* cur = h->xa;
  int ldcur = h->Ka;
  for (int i = 0; i < h->cfg.num_blocks; ++i) {
    const auto& w = h->bt[i];
    mark(h, ST_QKVT_GEMM, st);
    if (i == 0) {
      // QKVT = X0 @ W + b with the position / mark-code thirds of X0 folded (commit()): temporal.py:409
      GemmArgs g;
      g.A = h->xa; g.lda = h->Ka; g.W = h->wfold0T; g.ldw = h->Ka; g.w_is_nk = true; g.C = h->qkvt; g.ldc = 4 * d;
      g.Wlo = h->wfold0T + (size_t)h->Ka * 4 * d;
      g.M = (int)rows; g.N = 4 * d; g.K = h->Ka; g.pbias = h->pbias0; g.pperiod = L;
      if (f16 && (h->f16_mask & 1)) { g.W16 = h->mt16.at("wfold0"); g.a_amax = am + AMAX_XA; }
      EDGL_TRY(launch_gemm(g, st));
    } else {
      EDGL_TRY(dense_nk(cur, ldcur, h->btT[i].at("qkvt_w"), d, F(w, "qkvt_b"), h->qkvt, 4 * d, rows, 4 * d, ACT_NONE,
                        nullptr, 0, st, true, w16(i, "qkvt_w", 0), am + AMAX_LN2));
    }
    AttnArgs a = attn_args(h, w, h->qkvt, h->kmask, h->spans, h->marks, cur, ldcur, h->p0,
                           (tr && !tr->lam.empty()) ? tr->lam[i] : nullptr, B, false, true);
    bool att_amax = false;
    if (f16) { a.out_amax = am + AMAX_ATT; a.amax_published = &att_amax; }
    mark(h, ST_ATTENTION, st);
    EDGL_TRY(launch_attention(a, st));                                                   // temporal.py:412-447
    // a fallback attention kernel (shape not covered at run time) does not publish max|O|: take it in a pass of
    // its own rather than feed the scaled 3xFP16 attention-out GEMM a zero maximum
    if (f16 && (h->f16_mask & 2) && !att_amax) EDGL_TRY(launch_absmax(h->p0, rows * d, am + AMAX_ATT, st));
    if (h->ln_fuse && !tr) {
      // LayerNorm statistics from the producing GEMM's epilogue, applied by the consumers (LnEpi, common.cuh): no
      // LayerNorm pass reads or writes the activations
      const bool last = (i == h->cfg.num_blocks - 1);
      const int np = ln_stats_parts(d);
      LnEpi e_ao, e_ff1, e_ff2;
      e_ao.stats = h->ln_parts; e_ao.L = L;
      mark(h, ST_AO_GEMM, st);
      EDGL_TRY(dense_nk(h->p0, d, h->btT[i].at("ao_w"), d, F(w, "ao_b"), h->p1, d, rows, d, ACT_NONE, cur, ldcur, st, true,
                        nullptr, nullptr, nullptr, &e_ao));                                  // :113,116 (pre-LN)
      mark(h, ST_LN_ATT, st);
      EDGL_TRY(launch_ln_finalize(h->ln_parts, np, B, L, d, h->ln_rs1, nullptr, nullptr, nullptr, nullptr, nullptr, st));
      e_ff1.a_rs = h->ln_rs1; e_ff1.a_g = F(w, "ao_ln_g"); e_ff1.a_b = F(w, "ao_ln_b"); e_ff1.L = L;
      mark(h, ST_FF1_GEMM, st);
      EDGL_TRY(dense_nk(h->p1, d, h->btT[i].at("ff1_w"), d, F(w, "ff1_b"), h->p2, 2 * d, rows, 2 * d, ACT_GELU,
                        nullptr, 0, st, true, nullptr, nullptr, nullptr, &e_ff1));           // :116,120-121
      e_ff2.r_rs = h->ln_rs1; e_ff2.r_g = F(w, "ao_ln_g"); e_ff2.r_b = F(w, "ao_ln_b"); e_ff2.L = L;
      e_ff2.stats = h->ln_parts;
      mark(h, ST_FF2_GEMM, st);
      EDGL_TRY(dense_nk(h->p2, 2 * d, h->btT[i].at("ff2_w"), 2 * d, F(w, "ff2_b"), h->p0, d, rows, d, ACT_NONE, h->p1, d,
                        st, true, nullptr, nullptr, nullptr, &e_ff2));                       // :125,128 (pre-LN)
      mark(h, ST_LN_FF, st);
      EDGL_TRY(launch_ln_finalize(h->ln_parts, np, B, L, d, h->ln_rs2, nullptr, nullptr, nullptr, nullptr, nullptr, st));
      if (last) {
        // transform + its LayerNorm (EasyDGL.py:138-139): only the last row of a sequence is used (:146), so the layer
        // stores the last rows and the row statistics of all of them
        LnEpi e_tr;
        e_tr.a_rs = h->ln_rs2; e_tr.a_g = F(w, "ff_ln_g"); e_tr.a_b = F(w, "ff_ln_b"); e_tr.L = L;
        e_tr.stats = h->ln_parts; e_tr.last_only = 1;
        mark(h, ST_TR_GEMM, st);
        EDGL_TRY(dense_nk(h->p0, d, h->mtT.at("tr_w"), d, F(h->mt, "tr_b"), h->tr_last, d, rows, d, ACT_GELU, nullptr, 0, st,
                          true, nullptr, nullptr, nullptr, &e_tr));
        mark(h, ST_LN_OUT, st);
        EDGL_TRY(launch_ln_finalize(h->ln_parts, np, B, L, d, h->ln_rs3, h->tr_last, F(h->mt, "tr_ln_g"), F(h->mt, "tr_ln_b"),
                                    y, (f16 && y == h->y) ? am + AMAX_Y : nullptr, st, ldy));
        mark(h, ST_END, st);
        return 0;
      }
      // more blocks follow: the next block reads LayerNorm(p0) three times (QKVT, both residuals): materialise it
      EDGL_TRY(launch_layernorm(h->p0, F(w, "ff_ln_g"), F(w, "ff_ln_b"), B, L, d, h->p2, false, st,
                                f16 ? am + AMAX_LN2 : nullptr));
      cur = h->p2;
      ldcur = d;
      continue;
    }
    mark(h, ST_AO_GEMM, st);
    EDGL_TRY(dense_nk(h->p0, d, h->btT[i].at("ao_w"), d, F(w, "ao_b"), h->p1, d, rows, d, ACT_NONE, cur, ldcur, st, true,
                      w16(i, "ao_w", 1), am + AMAX_ATT));                                   // :113,116
    mark(h, ST_LN_ATT, st);
    EDGL_TRY(launch_layernorm(h->p1, F(w, "ao_ln_g"), F(w, "ao_ln_b"), B, L, d, h->p0, false, st,
                              f16 ? am + AMAX_LN1 : nullptr));                           // :116
    mark(h, ST_FF1_GEMM, st);
    EDGL_TRY(dense_nk(h->p0, d, h->btT[i].at("ff1_w"), d, F(w, "ff1_b"), h->p2, 2 * d, rows, 2 * d, ACT_GELU, nullptr,
                      0, st, true, w16(i, "ff1_w", 2), am + AMAX_LN1, f16 ? am + AMAX_FF1 : nullptr));  // :120-121
    mark(h, ST_FF2_GEMM, st);
    EDGL_TRY(dense_nk(h->p2, 2 * d, h->btT[i].at("ff2_w"), 2 * d, F(w, "ff2_b"), h->p1, d, rows, d, ACT_NONE, h->p0, d,
                      st, true, w16(i, "ff2_w", 3), am + AMAX_FF1));                        // :125,128
    mark(h, ST_LN_FF, st);
    EDGL_TRY(launch_layernorm(h->p1, F(w, "ff_ln_g"), F(w, "ff_ln_b"), B, L, d, h->p2, false, st,
                              f16 ? am + AMAX_LN2 : nullptr));                           // :128
    cur = h->p2;
    ldcur = d;
  }
  mark(h, ST_TR_GEMM, st);
  EDGL_TRY(dense_nk(cur, ldcur, h->mtT.at("tr_w"), d, F(h->mt, "tr_b"), h->p0, d, rows, d, ACT_GELU, nullptr, 0, st, true,
                    (f16 && (h->f16_mask & 16)) ? h->mt16.at("tr_w") : nullptr, am + AMAX_LN2));  // :138
  mark(h, ST_LN_OUT, st);
  if (tr) {  // every position is needed (EasyDGL.py:141: batch_gather at the masked positions)
    EDGL_TRY(launch_layernorm(h->p0, F(h->mt, "tr_ln_g"), F(h->mt, "tr_ln_b"), B, L, d, tr->Y, false, st));
    mark(h, ST_END, st);
    return 0;
  }
  EDGL_TRY(launch_layernorm(h->p0, F(h->mt, "tr_ln_g"), F(h->mt, "tr_ln_b"), B, L, d, y, true, st,
                            (f16 && y == h->y) ? am + AMAX_Y : nullptr, ldy));            // :139,146
  mark(h, ST_END, st);
  return 0;
}

// CTSMA.__call__ up to y (CTSMA.py:46-87)
int encode_ctsma(edgl_handle* h, const int64_t* ids, const float* ts, int B, float* y, long long ldy, cudaStream_t st,
                 const TrainOut* tr = nullptr) {
  const int d = h->d, L = h->L;
  const long long rows = (long long)B * L;
  EmbedArgs e = embed_args(h, ids, ts, B);
  e.X0 = h->p2; e.ldx0 = 2 * d; e.spans = h->spans; e.marks = h->marks; e.kmask = h->kmask;
  mark(h, ST_EMBED, st);
  EDGL_TRY(launch_embed(e, st));
  float* cur = h->p2;
  int cin = 2 * d;
  for (int i = 0; i < h->cfg.num_blocks; ++i) {
    const auto& w = h->bt[i];
    mark(h, ST_LN_IN, st);
    EDGL_TRY(launch_layernorm(cur, F(w, "ln1_g"), F(w, "ln1_b"), B, L, cin, h->p0, false, st));  // CTSMA.py:68
    mark(h, ST_QKVT_GEMM, st);
    EDGL_TRY(dense_nk(h->p0, cin, h->btT[i].at("q_w"), cin, F(w, "q_b"), h->qkvt, 4 * d, rows, d, ACT_NONE, nullptr, 0,
                      st));
    EDGL_TRY(dense_nk(cur, cin, h->btT[i].at("kvt_w"), cin, h->bkvt[i], h->qkvt + d, 4 * d, rows, 3 * d, ACT_NONE,
                      nullptr, 0, st));                                                  // temporal.py:340-343
    AttnArgs a = attn_args(h, w, h->qkvt, h->kmask, h->spans, h->marks, h->p0, cin, h->p1,
                           (tr && !tr->lam.empty()) ? tr->lam[i] : nullptr, B, true, false);
    mark(h, ST_ATTENTION, st);
    EDGL_TRY(launch_attention(a, st));                                                   // temporal.py:345-385
    mark(h, ST_LN_ATT, st);
    EDGL_TRY(launch_layernorm(h->p1, F(w, "ln2_g"), F(w, "ln2_b"), B, L, d, h->p0, false, st));  // CTSMA.py:73
    mark(h, ST_FF1_GEMM, st);
    EDGL_TRY(dense_nk(h->p0, d, h->btT[i].at("ff1_w"), d, F(w, "ff1_b"), h->p1, d, rows, d, ACT_RELU, nullptr, 0, st));  // Base.py:79
    mark(h, ST_FF2_GEMM, st);
    EDGL_TRY(dense_nk(h->p1, d, h->btT[i].at("ff2_w"), d, F(w, "ff2_b"), h->p2, d, rows, d, ACT_NONE, h->p0, d, st));    // Base.py:83,86
    cur = h->p2;
    cin = d;
  }
  mark(h, ST_LN_OUT, st);
  if (tr) {  // CTSMA.py:80,83: every position is predicted in training
    EDGL_TRY(launch_layernorm(cur, F(h->mt, "out_ln_g"), F(h->mt, "out_ln_b"), B, L, d, tr->Y, false, st));
    mark(h, ST_END, st);
    return 0;
  }
  EDGL_TRY(launch_layernorm(cur, F(h->mt, "out_ln_g"), F(h->mt, "out_ln_b"), B, L, d, y, true, st, nullptr, ldy));  // CTSMA.py:80,87
  mark(h, ST_END, st);
  return 0;
}

int encode(edgl_handle* h, const int64_t* ids, const float* ts, int B, float* y, cudaStream_t st, long long ldy = 0,
           const TrainOut* tr = nullptr) {
  if (B == 0) return 0;
  if (ldy == 0) ldy = h->d;
  return h->cfg.model == EDGL_MODEL_EASYDGL ? encode_easydgl(h, ids, ts, B, y, ldy, st, tr)
                                            : encode_ctsma(h, ids, ts, B, y, ldy, st, tr);
}

// Encoder in training mode + the hidden rows that are predicted: Yg [B*M, d] (gathered at `positions`, or every row
// when positions == null and M == L).  Grows the handle's training buffers on demand.
int train_encode(edgl_handle* h, const int64_t* ids, const float* ts, int B, const int64_t* positions, int M,
                 bool want_lam, const float** Yg, cudaStream_t st) {
  const long long rows = (long long)B * M;
  const int nb = h->cfg.num_blocks;
  if (rows > h->train_rows_cap) {
    const long long cap = (long long)h->cfg.max_batch * (M > h->L ? M : h->L);
    float* p = nullptr;
    EDGL_TRY(dev_alloc(h, &p, (size_t)cap * h->d)); h->train_y = p;
    EDGL_TRY(dev_alloc(h, &p, (size_t)cap)); h->train_pe = p;
    EDGL_TRY(dev_alloc(h, &p, (size_t)cap)); h->train_wt = p;
    EDGL_TRY(dev_alloc(h, &p, (size_t)3 * h->h * cap)); h->train_tpp = p;
    h->train_rows_cap = cap;
  }
  if (!h->train_acc) EDGL_TRY(dev_alloc(h, &h->train_acc, (size_t)(3 + 3 * nb)));
  TrainOut tr;
  tr.Y = h->p1;  // free after the last dense layer of either model
  if (want_lam) {
    if ((int)h->train_lam.size() != nb) h->train_lam.assign(nb, nullptr);
    for (int i = 0; i < nb; ++i)
      if (!h->train_lam[i]) EDGL_TRY(dev_alloc(h, &h->train_lam[i], (size_t)h->h * h->cfg.max_batch * h->L * h->E));
    tr.lam = h->train_lam;
  }
  if (h->cfg.model == EDGL_MODEL_CTSMA) tr.Y = h->p0;  // CTSMA's last block leaves its output in p2
  EDGL_TRY(encode(h, ids, ts, B, nullptr, st, 0, &tr));
  if (positions) {
    EDGL_CUDA(cudaMemsetAsync(h->flag, 0, sizeof(int), st));
    EDGL_TRY(launch_gather_rows(tr.Y, positions, h->L, M, h->d, rows, h->train_y, h->flag, st));
    *Yg = h->train_y;
  } else {
    *Yg = tr.Y;
  }
  return 0;
}

// logits[r0:r0+rc, c0:c1] = y @ table[c0:c1]^T + bias   (EasyDGL.py:149-150 / CTSMA.py:89-90, Base.py:106-110)
// ncols > 0: only the first ncols columns of the shard; flt: candidate filter instead of the store (tensor-core kernel);
// run_if: predicated launch
bool logits_use_f16(const edgl_handle* h, const float* y) {
  return h->f16_gemm && (h->f16_mask & 32) && y == h->y && h->mt16.count("table");
}
int logits_rows(edgl_handle* h, const float* y, int ldy, long long rc, float* out, int ldo, cudaStream_t st,
                int ncols = 0, const TopkFilter* flt = nullptr, const int* run_if = nullptr) {
  GemmArgs g;
  g.A = y; g.lda = ldy;
  g.W = F(h->mt, "item_embs") + h->c0 * h->d; g.ldw = h->d; g.w_is_nk = true;
  g.Wlo = h->table_lo;
  if (logits_use_f16(h, y) && !flt && !run_if) {  // the local encoder's y: its maximum is in AMAX_Y
    g.W16 = h->mt16.at("table");
    g.a_amax = h->amax + AMAX_Y;
  }
  g.zero_wrow0 = (h->c0 == 0);  // zero_pad=True: row 0 of the tied table is zeros (coding.py:56-57)
  g.C = out; g.ldc = ldo; g.M = (int)rc; g.N = ncols > 0 ? ncols : (int)(h->c1 - h->c0); g.K = h->d;
  g.bias = h->bias_full + h->c0;
  if (flt) g.flt = *flt;
  g.run_if = run_if;
  return launch_gemm(g, st);
}

// Fused logits + seen-mask + top-K of one chunk of rows WITHOUT materialising [rows, Ns] (EasyDGL.py:149-150,
// Base.py:156-181; SURVEY 7.6 / 8e "top-K in the GEMM epilogue"):
//   1. logits of a column SAMPLE (the first ns columns of the shard), seen-mask, its K-th largest value per row =
//      thr[row]: at least K unmasked logits of the row are >= thr, so thr is a lower bound of the row's K-th largest;
//   2. the logits GEMM over all columns with the candidate-filter epilogue: (value, column) of everything >= thr;
//   3. exact top-K of the unmasked candidates (topk_select_kernel).
// Same kernel and same accumulation order as the materialised path, so the selected values are bit-identical to it.
// A row with too many candidates (ties, a constant row, a sample that under-estimates) or too few raises a device
// flag and the materialised path - launched unconditionally, predicated on that flag - recomputes the chunk.
bool topk_fused_ok(const edgl_handle* h, const float* y, long long Ns, long long ldw, int* ns_out, int* cap_out) {
  // Measured (one B200, B = 4096 / 1024 / 2048 rows): C5's 1 000 001 columns 12.1 -> 7.9 ms (the [B, N1] round trip
  // and the 4 re-streams of the table are gone), C4's 100 001 columns 0.42 -> 0.46 ms and C2's 18 001 columns
  // 0.27 -> 0.46 ms (two more per-row kernels and seven more launches than the materialised path: fixed costs that
  // short rows do not amortise).  So the fused path is the default for shards of >= 200 000 columns;
  // EDGL_TOPK_FUSE=1 forces it for every shard it supports (>= 8192 columns), =0 turns it off;
  // EDGL_TOPK_SAMPLE=n sets the sample width.  Read per call: test switches.
  const char* fe = getenv("EDGL_TOPK_FUSE");
  const bool off = (fe && fe[0] == '0') || (!(fe && fe[0] == '1') && Ns < 200000);
  const char* se = getenv("EDGL_TOPK_SAMPLE");
  const int ns_env = se ? atoi(se) : 0;
  static const bool simt = [] { const char* e = getenv("EDGL_GEMM"); return e && e[0] == 's'; }();
  if (off || simt || logits_use_f16(h, y) || h->K > 256 || (h->d % 4) != 0) return false;
  long long ns = ns_env > 0 ? ns_env : ((Ns / 16 + 255) / 256) * 256;
  if (ns < 2048) ns = 2048;
  if (ns * 4 > Ns || ns < 8 * h->K) return false;  // short rows: the sample would be most of the row
  int cap = 4096;
  if (ns + 2 * cap + 1 + 2 * h->K > ldw) cap = 2048;
  if (ns + 2 * cap + 1 + 2 * h->K > ldw) return false;
  *ns_out = (int)ns;
  *cap_out = cap;
  return true;
}

int logits_topk(edgl_handle* h, const float* y, int ldy, const int64_t* seen, int seen_len, long long seen_stride,
                long long Bt, int32_t* idx, float* val, long long ostride, cudaStream_t st,
                const TopkP2P* p2p = nullptr) {
  if (ostride == 0) ostride = h->K;
  if (ldy == 0) ldy = h->d;
  if (seen_stride == 0) seen_stride = seen_len;
  const int Ns = (int)(h->c1 - h->c0);
  const int ldw = (Ns + 3) & ~3;  // padded pitch: vector stores in the GEMM epilogue
  int ns = 0, cap = 0;
  const bool fused = !p2p && topk_fused_ok(h, y, Ns, ldw, &ns, &cap);
  for (long long r0 = 0; r0 < Bt; r0 += h->ws_rows) {
    const long long rc = (Bt - r0 < h->ws_rows) ? (Bt - r0) : h->ws_rows;
    const int64_t* seen_r = seen ? seen + r0 * seen_stride : nullptr;
    const int* run_if = nullptr;
    if (fused) {
      // workspace of the chunk, carved out of logits_ws: sample logits | candidates | counters | sample top-K
      float* samp = h->logits_ws;
      unsigned long long* cand = reinterpret_cast<unsigned long long*>(samp + rc * ns);
      unsigned int* cnt = reinterpret_cast<unsigned int*>(cand + rc * cap);
      float* tval = reinterpret_cast<float*>(cnt + rc);
      int32_t* tidx = reinterpret_cast<int32_t*>(tval + rc * h->K);
      mark(h, ST_LOGITS_GEMM, st);
      EDGL_TRY(logits_rows(h, y + r0 * ldy, ldy, rc, samp, ns, st, ns));
      if (seen) {
        mark(h, ST_MASK_SEEN, st);
        EDGL_TRY(launch_mask_seen(samp, ns, (int)rc, seen_r, seen_len, seen_stride, h->c0, h->c0 + ns, st));
      }
      mark(h, ST_TOPK, st);
      EDGL_TRY(launch_topk(samp, ns, (int)rc, ns, h->K, 0, h->K, tidx, tval, st));
      EDGL_CUDA(cudaMemsetAsync(cnt, 0, (size_t)rc * sizeof(unsigned int), st));
      EDGL_CUDA(cudaMemsetAsync(h->topk_redo, 0, sizeof(int), st));
      mark(h, ST_LOGITS_GEMM, st);
      TopkFilter flt;
      flt.thr = tval + (h->K - 1); flt.thr_stride = h->K;
      flt.cand = cand; flt.cap = cap; flt.cnt = cnt;
      EDGL_TRY(logits_rows(h, y + r0 * ldy, ldy, rc, nullptr, 0, st, 0, &flt));
      mark(h, ST_TOPK, st);
      EDGL_TRY(launch_topk_select(cand, cnt, cap, (int)rc, h->K, (int)h->c0, seen_r, seen_len, seen_stride, h->c0, h->c1,
                                  ostride, idx + r0 * ostride, val + r0 * ostride, h->topk_redo, st));
      run_if = h->topk_redo;  // the launches below do nothing unless a row asked for the materialised path
    }
    mark(h, ST_LOGITS_GEMM, st);
    EDGL_TRY(logits_rows(h, y + r0 * ldy, ldy, rc, h->logits_ws, ldw, st, 0, nullptr, run_if));
    if (seen) {
      mark(h, ST_MASK_SEEN, st);
      EDGL_TRY(launch_mask_seen(h->logits_ws, ldw, (int)rc, seen_r, seen_len, seen_stride, h->c0, h->c1, st, run_if));
    }
    mark(h, ST_TOPK, st);
    if (p2p) {
      TopkP2P pp = *p2p;
      pp.row_base = (int)r0;
      pp.signal = (r0 + rc >= Bt) ? 1 : 0;  // raise the peers' flags after the last chunk only
      EDGL_TRY(launch_topk(h->logits_ws, ldw, (int)rc, Ns, h->K, (int)h->c0, 0, nullptr, nullptr, st, &pp));
    } else {
      EDGL_TRY(launch_topk(h->logits_ws, ldw, (int)rc, Ns, h->K, (int)h->c0, ostride, idx + r0 * ostride,
                           val + r0 * ostride, st, nullptr, run_if));
    }
  }
  mark(h, ST_END, st);
  return 0;
}

}  // namespace

extern "C" {

const char* edgl_last_error(void) { return last_error().c_str(); }
int edgl_version(void) { return 100; }
int64_t edgl_launch_count(void) { return (int64_t)g_launches.load(); }

int edgl_create(const edgl_config* cfg, edgl_handle** out) {
  if (!cfg || !out) return set_error(EDGL_EINVAL, "null argument");
  *out = nullptr;
  EDGL_REQUIRE(cfg->model == EDGL_MODEL_EASYDGL || cfg->model == EDGL_MODEL_CTSMA,
               "The ranking model: %d not implemented", cfg->model);  // util.py:96
  EDGL_REQUIRE(cfg->max_batch >= 1 && cfg->seq_len >= 1 && cfg->num_units >= 4 && cfg->num_heads >= 1 &&
                   cfg->num_blocks >= 1 && cfg->num_events >= 1 && cfg->num_rows >= 2 && cfg->mark_rows >= 1 &&
                   cfg->topk >= 1,
               "edgl_create: non-positive size in config");
  EDGL_REQUIRE(cfg->num_units % cfg->num_heads == 0, "num_units %d not divisible by num_heads %d", cfg->num_units,
               cfg->num_heads);
  EDGL_REQUIRE(cfg->num_units % 4 == 0, "num_units must be a multiple of 4 (got %d)", cfg->num_units);
  EDGL_REQUIRE(cfg->shard_world >= 1 && cfg->shard_rank >= 0 && cfg->shard_rank < cfg->shard_world,
               "bad shard rank/world %d/%d", cfg->shard_rank, cfg->shard_world);
  EDGL_REQUIRE(cfg->time_scale > 0.f, "time_scale must be positive");
  int dev = 0;
  EDGL_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  EDGL_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10)
    return set_error(EDGL_EARCH, "device %d is sm_%d%d; this library is built for sm_100a only", dev, prop.major,
                     prop.minor);
  edgl_handle* h = new edgl_handle();
  h->cfg = *cfg;
  h->dev = dev;
  h->L = cfg->seq_len; h->d = cfg->num_units; h->h = cfg->num_heads; h->dh = h->d / h->h; h->E = cfg->num_events;
  h->N1 = cfg->num_rows; h->K = cfg->topk;
  const bool easy = cfg->model == EDGL_MODEL_EASYDGL;
  h->ts_len = easy ? h->L : h->L + 1;
  h->Ka = easy ? h->d + h->E : 2 * h->d;
  const long long per = (h->N1 + cfg->shard_world - 1) / cfg->shard_world;
  h->c0 = per * cfg->shard_rank;
  h->c1 = h->c0 + per < h->N1 ? h->c0 + per : h->N1;
  if (h->c0 > h->c1) h->c0 = h->c1;
  h->bt.resize(cfg->num_blocks);
  h->btT.resize(cfg->num_blocks);
  h->bt16.resize(cfg->num_blocks);
  h->wkvt.assign(cfg->num_blocks, nullptr);
  h->bkvt.assign(cfg->num_blocks, nullptr);
  const long long rows = (long long)cfg->max_batch * h->L;
  const int d = h->d;
  int rc = 0;
#define EDGL_ALLOC(p, n)                         \
  if (!rc) rc = dev_alloc(h, &(p), (size_t)(n));
  EDGL_ALLOC(h->tscale, d / 2);
  EDGL_ALLOC(h->mark8, (long long)cfg->mark_rows * h->E);
  EDGL_ALLOC(h->flag, 1);
  EDGL_ALLOC(h->topk_redo, 1);
  EDGL_ALLOC(h->p2p_counter, 4);
  if (!rc && cudaMemset(h->p2p_counter, 0, 4 * sizeof(unsigned int)) != cudaSuccess) rc = set_error(EDGL_ECUDA, "cudaMemset failed");
  EDGL_ALLOC(h->bias_full, h->N1);
  if (easy) {
    EDGL_ALLOC(h->wfold0, (long long)h->Ka * 4 * d);
    EDGL_ALLOC(h->pbias0, (long long)h->L * 4 * d);
    EDGL_ALLOC(h->xa, rows * h->Ka);
  } else {
    for (int i = 0; i < cfg->num_blocks; ++i) {
      EDGL_ALLOC(h->wkvt[i], (long long)cin_of(h, i) * 3 * d);
      EDGL_ALLOC(h->bkvt[i], 3 * d);
    }
  }
  EDGL_ALLOC(h->p0, rows * 2 * d);
  EDGL_ALLOC(h->p1, rows * 2 * d);
  EDGL_ALLOC(h->p2, rows * 2 * d);
  EDGL_ALLOC(h->qkvt, rows * 4 * d);
  EDGL_ALLOC(h->spans, rows);
  EDGL_ALLOC(h->marks, rows * h->E);
  EDGL_ALLOC(h->kmask, rows);
  EDGL_ALLOC(h->y, (long long)cfg->max_batch * d);
  EDGL_ALLOC(h->amax, AMAX_COUNT);
  {
    // LayerNorm folded into the tensor-core dense layers: EasyDGL, d a multiple of 16, tensor-core GEMMs in use
    const char* ge = getenv("EDGL_GEMM");
    const char* le = getenv("EDGL_LN_FUSE");
    h->ln_fuse = easy && d % 16 == 0 && d <= 256 && !(ge && ge[0] == 's') && !(le && le[0] == '0');
    if (h->ln_fuse) {
      EDGL_ALLOC(h->ln_parts, rows * ln_stats_parts(d));
      EDGL_ALLOC(h->ln_rs1, cfg->max_batch);
      EDGL_ALLOC(h->ln_rs2, cfg->max_batch);
      EDGL_ALLOC(h->ln_rs3, cfg->max_batch);
      EDGL_ALLOC(h->tr_last, (long long)cfg->max_batch * d);
    }
  }
  {
    // The scaled 3xFP16 dense layers (gemm_f16.cu) need every activation's running maximum from its producer, and
    // the attention producer that publishes one is attn_f16.cu: EasyDGL with dh = 16, E = 16, L <= 208 and the default
    // attention kernel.  By default only the block-0 QKVT layer takes it (f16_mask bit 0): it is the one dense layer
    // bound by the shared-memory port (0.40 -> 0.35 ms at C2); attention-out / FF2 already run at 60-72 % of the HBM
    // copy rate and FF1 / transform are bound by the erf of their GELU epilogue, so fp16 operands change nothing
    // there (measured).  EDGL_F16_MASK=63 puts all six on it (parity-tested); EDGL_GEMM=tf32 none.
    const char* ge = getenv("EDGL_GEMM");
    const char* ae = getenv("EDGL_ATTN");
    h->f16_gemm = easy && attention_f16_pack_bytes(h->dh, h->E) != 0 && h->L <= 208 && (!ae || ae[0] == 'f' || ae[0] == 'd' || (ae[0] == 't' && ae[1] == 'c' && ae[2] == '2')) &&
                  !(ge && (ge[0] == 't' || ge[0] == 's'));
    if (const char* me = getenv("EDGL_F16_MASK")) h->f16_mask = atoi(me);
  }
  {
    const long long Ns = h->c1 - h->c0 > 0 ? h->c1 - h->c0 : 1;
    const long long max_bt = (long long)cfg->max_batch * cfg->shard_world;
    const long long ldw = (Ns + 3) & ~3ll;
    long long r = (2ll << 30) / (ldw * 4);
    if (r < 1) r = 1;
    if (r > max_bt) r = max_bt;
    h->ws_rows = r;
    EDGL_ALLOC(h->logits_ws, r * ldw);
  }
  for (int sl = 0; sl < 2; ++sl) {
    EDGL_ALLOC(h->st_ids[sl], rows);
    EDGL_ALLOC(h->st_ts[sl], (long long)cfg->max_batch * h->ts_len);
    EDGL_ALLOC(h->st_idx[sl], (long long)cfg->max_batch * h->K);
    EDGL_ALLOC(h->st_val[sl], (long long)cfg->max_batch * h->K);
  }
#undef EDGL_ALLOC
  if (!rc) {
    bool ok = cudaStreamCreateWithFlags(&h->cs_in, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&h->cs_out, cudaStreamNonBlocking) == cudaSuccess;
    for (int sl = 0; sl < 2 && ok; ++sl)
      ok = cudaEventCreateWithFlags(&h->ev_h2d[sl], cudaEventDisableTiming) == cudaSuccess &&
           cudaEventCreateWithFlags(&h->ev_done[sl], cudaEventDisableTiming) == cudaSuccess &&
           cudaEventCreateWithFlags(&h->ev_d2h[sl], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) rc = set_error(EDGL_ECUDA, "could not create the copy streams / events of the host entry point");
  }
  if (rc) {
    edgl_destroy(h);
    return rc;
  }
  *out = h;
  return 0;
}

int edgl_destroy(edgl_handle* h) {
  if (!h) return 0;
  for (void* p : h->owned) cudaFree(p);
  for (cudaEvent_t e : h->prof_ev) cudaEventDestroy(e);
  for (int sl = 0; sl < 2; ++sl) {
    if (h->ev_h2d[sl]) cudaEventDestroy(h->ev_h2d[sl]);
    if (h->ev_done[sl]) cudaEventDestroy(h->ev_done[sl]);
    if (h->ev_d2h[sl]) cudaEventDestroy(h->ev_d2h[sl]);
  }
  if (h->cs_in) cudaStreamDestroy(h->cs_in);
  if (h->cs_out) cudaStreamDestroy(h->cs_out);
  delete h;
  return 0;
}

int edgl_get_config(const edgl_handle* h, edgl_config* out) {
  if (!h || !out) return set_error(EDGL_EINVAL, "null argument");
  *out = h->cfg;
  return 0;
}

int edgl_num_stages(void) { return ST_COUNT - 1; }
const char* edgl_stage_name(int stage) { return (stage >= 0 && stage < ST_COUNT) ? kStageNames[stage] : ""; }

int edgl_profile(edgl_handle* h, int enable) {
  if (!h) return set_error(EDGL_EINVAL, "null handle");
  if (enable && h->prof_ev.empty()) {
    h->prof_ev.resize(32768);
    h->prof_stage.assign(32768, ST_END);
    for (auto& e : h->prof_ev) EDGL_CUDA(cudaEventCreate(&e));
  }
  h->prof_on = enable != 0;
  h->prof_n = 0;
  return 0;
}

int edgl_profile_read(edgl_handle* h, double* ms, int64_t* count, int n) {
  if (!h || !ms || !count) return set_error(EDGL_EINVAL, "null argument");
  EDGL_REQUIRE(n >= ST_COUNT - 1, "edgl_profile_read: need room for %d stages", ST_COUNT - 1);
  for (int i = 0; i < n; ++i) { ms[i] = 0.0; count[i] = 0; }
  if (h->prof_n == 0) return 0;
  EDGL_CUDA(cudaEventSynchronize(h->prof_ev[h->prof_n - 1]));
  for (size_t i = 0; i + 1 < h->prof_n; ++i) {
    const int s = h->prof_stage[i];
    if (s == ST_END) continue;
    float t = 0.f;
    EDGL_CUDA(cudaEventElapsedTime(&t, h->prof_ev[i], h->prof_ev[i + 1]));
    ms[s] += t;
    count[s] += 1;
  }
  h->prof_n = 0;
  return 0;
}

int edgl_set_tensor(edgl_handle* h, const char* name, int block, const void* dev_ptr, int64_t numel) {
  if (!h || !name || !dev_ptr) return set_error(EDGL_EINVAL, "null argument");
  EDGL_REQUIRE(block >= -1 && block < h->cfg.num_blocks, "block %d out of range", block);
  const long long want = expected_numel(h, name, block);
  EDGL_REQUIRE(want >= 0, "unknown tensor name '%s' (block %d)", name, block);
  EDGL_REQUIRE(want == numel, "tensor '%s' (block %d): expected %lld elements, got %lld", name, block, want,
               (long long)numel);
  EDGL_REQUIRE((reinterpret_cast<uintptr_t>(dev_ptr) & 15) == 0, "tensor '%s' must be 16-byte aligned", name);
  Tensor t;
  t.p = dev_ptr;
  t.numel = numel;
  if (block < 0) h->mt[name] = t; else h->bt[block][name] = t;
  h->committed = false;
  return 0;
}

int edgl_commit(edgl_handle* h, void* stream) {
  if (!h) return set_error(EDGL_EINVAL, "null handle");
  cudaStream_t st = (cudaStream_t)stream;
  const bool easy = h->cfg.model == EDGL_MODEL_EASYDGL;
  const int d = h->d, E = h->E, L = h->L;
  // every variable of the reference graph must be bound
  if (easy) {
    for (const char* n : kModelNamesEasy)
      if (!h->mt.count(n)) return set_error(EDGL_ESTATE, "tensor '%s' not set", n);
    for (int i = 0; i < h->cfg.num_blocks; ++i)
      for (const char* n : kBlockNamesEasy)
        if (!h->bt[i].count(n)) return set_error(EDGL_ESTATE, "tensor '%s' of block %d not set", n, i);
  } else {
    for (const char* n : kModelNamesCtsma)
      if (!h->mt.count(n)) return set_error(EDGL_ESTATE, "tensor '%s' not set", n);
    for (int i = 0; i < h->cfg.num_blocks; ++i)
      for (const char* n : kBlockNamesCtsma)
        if (!h->bt[i].count(n)) return set_error(EDGL_ESTATE, "tensor '%s' of block %d not set", n, i);
  }
  // TimeSinusoidCoding.__init__ (coding.py:134-135): float64 power, stored as fp32
  {
    std::vector<float> sc(d / 2);
    for (int j = 0; j < d / 2; ++j) sc[j] = (float)pow(10000.0, (double)(2 * j) * 1.0 / (double)d);
    EDGL_CUDA(cudaMemcpyAsync(h->tscale, sc.data(), sc.size() * sizeof(float), cudaMemcpyHostToDevice, st));
    EDGL_CUDA(cudaStreamSynchronize(st));  // sc goes out of scope
  }
  // mark_lookup_table (EasyDGL.py:45): int64 -> uint8, values must index mark_embs (EasyDGL.py:87)
  EDGL_CUDA(cudaMemsetAsync(h->flag, 0, sizeof(int), st));
  EDGL_TRY(launch_mark_table_to_u8(reinterpret_cast<const int64_t*>(h->mt.at("mark_table").p),
                                   (long long)h->cfg.mark_rows * E, h->mark8, h->flag, E, st));
  // output_bias(inf_pad=True) (Base.py:106-110)
  {
    const float m1000 = -1000.f;
    EDGL_CUDA(cudaMemcpyAsync(h->bias_full, &m1000, sizeof(float), cudaMemcpyHostToDevice, st));
    EDGL_CUDA(cudaMemcpyAsync(h->bias_full + 1, F(h->mt, "output_bias"), (size_t)(h->N1 - 1) * sizeof(float),
                              cudaMemcpyDeviceToDevice, st));
  }
  if (easy) {
    // Block-0 input is X0 = [x | pos | mcode] (EasyDGL.py:89); pos depends only on l and mcode is
    // linear in the mark-value histogram, so  X0 @ W = x @ W[0:d] + (pos @ W[d:2d])[l] + cnt @ (mark_embs_zp @ W[2d:3d]).
    const auto& w = h->bt[0];
    const float* W = F(w, "qkvt_w");
    EDGL_CUDA(cudaMemcpyAsync(h->wfold0, W, (size_t)d * 4 * d * sizeof(float), cudaMemcpyDeviceToDevice, st));
    EDGL_TRY(dense(F(h->mt, "mark_embs"), d, W + (size_t)2 * d * 4 * d, 4 * d, nullptr, h->wfold0 + (size_t)d * 4 * d,
                   4 * d, E, 4 * d, d, ACT_NONE, nullptr, 0, st));
    EDGL_CUDA(cudaMemsetAsync(h->wfold0 + (size_t)d * 4 * d, 0, (size_t)4 * d * sizeof(float), st));  // zero_pad row 0
    EDGL_TRY(dense(F(h->mt, "pos_embs"), d, W + (size_t)d * 4 * d, 4 * d, F(w, "qkvt_b"), h->pbias0, 4 * d, L, 4 * d, d,
                   ACT_NONE, nullptr, 0, st));
  } else {
    for (int i = 0; i < h->cfg.num_blocks; ++i) {
      const auto& w = h->bt[i];
      const int cin = cin_of(h, i);
      const char* wn[3] = {"k_w", "v_w", "t_w"};
      const char* bn[3] = {"k_b", "v_b", "t_b"};
      for (int j = 0; j < 3; ++j) {
        EDGL_CUDA(cudaMemcpy2DAsync(h->wkvt[i] + j * d, (size_t)3 * d * sizeof(float), F(w, wn[j]),
                                    (size_t)d * sizeof(float), (size_t)d * sizeof(float), cin,
                                    cudaMemcpyDeviceToDevice, st));
        EDGL_CUDA(cudaMemcpyAsync(h->bkvt[i] + j * d, F(w, bn[j]), (size_t)d * sizeof(float),
                                  cudaMemcpyDeviceToDevice, st));
      }
    }
  }
  // intensity-MLP constants in the fragment layout of the 3xFP16 attention kernel (attn_f16.cu)
  if (const size_t pb = attention_f16_pack_bytes(h->dh, E)) {
    h->mlp_pack.resize(h->cfg.num_blocks, nullptr);
    for (int i = 0; i < h->cfg.num_blocks; ++i) {
      const auto& w = h->bt[i];
      if (!h->mlp_pack[i]) EDGL_TRY(dev_alloc(h, &h->mlp_pack[i], pb));
      EDGL_TRY(launch_attention_f16_pack(F(w, "int_w"), F(w, "int_b"), F(w, "int_weight"), F(w, "int_scaling"), h->dh,
                                         E, h->mlp_pack[i], st));
    }
  }
  if (const size_t pb = attention_tc2_pack_bytes(h->dh, E)) {
    h->mlp_pack2.resize(h->cfg.num_blocks, nullptr);
    for (int i = 0; i < h->cfg.num_blocks; ++i) {
      const auto& w = h->bt[i];
      if (!h->mlp_pack2[i]) EDGL_TRY(dev_alloc(h, &h->mlp_pack2[i], pb));
      EDGL_TRY(launch_attention_tc2_pack(F(w, "int_w"), F(w, "int_b"), F(w, "int_weight"), F(w, "int_scaling"), h->dh,
                                         E, h->mlp_pack2[i], st));
    }
  }
  // tf32 lo part of the owned rows of the tied item table (B operand of the logits GEMM); skipped above 2 GiB
  {
    const long long n = (h->c1 - h->c0) * (long long)d;
    if (n * 4 <= (2ll << 30)) {
      if (!h->table_lo) EDGL_TRY(dev_alloc(h, &h->table_lo, (size_t)n));
      EDGL_TRY(launch_tf32_lo(F(h->mt, "item_embs") + h->c0 * d, n, h->table_lo, st));
    }
  }
  // K-major copies of the dense kernels for the tcgen05 GEMM (allocated once, refreshed on every commit)
  {
    auto transposed = [&](std::map<std::string, float*>& dst, const std::string& key, const float* src, int K,
                          int N) -> int {
      float*& buf = dst[key];
      if (!buf) EDGL_TRY(dev_alloc(h, &buf, (size_t)2 * K * N));  // [N,K] copy, then its tf32 lo part (gemm_tc.cu)
      EDGL_TRY(launch_transpose(src, K, N, buf, st));
      return launch_tf32_lo(buf, (long long)K * N, buf + (size_t)K * N, st);
    };
    for (int i = 0; i < h->cfg.num_blocks; ++i) {
      const auto& w = h->bt[i];
      const int cin = cin_of(h, i);
      if (easy) {
        // block 0's raw [3d,4d] kernel is only used by the layer-level entry point (the pipeline runs the folded one)
        EDGL_TRY(transposed(h->btT[i], i > 0 ? "qkvt_w" : "qkvt_w_raw", F(w, "qkvt_w"), cin, 4 * d));
        EDGL_TRY(transposed(h->btT[i], "ao_w", F(w, "ao_w"), d, d));
        EDGL_TRY(transposed(h->btT[i], "ff1_w", F(w, "ff1_w"), d, 2 * d));
        EDGL_TRY(transposed(h->btT[i], "ff2_w", F(w, "ff2_w"), 2 * d, d));
      } else {
        EDGL_TRY(transposed(h->btT[i], "q_w", F(w, "q_w"), cin, d));
        EDGL_TRY(transposed(h->btT[i], "kvt_w", h->wkvt[i], cin, 3 * d));
        EDGL_TRY(transposed(h->btT[i], "ff1_w", F(w, "ff1_w"), d, d));
        EDGL_TRY(transposed(h->btT[i], "ff2_w", F(w, "ff2_w"), d, d));
      }
    }
    if (easy) {
      EDGL_TRY(transposed(h->mtT, "tr_w", F(h->mt, "tr_w"), d, d));
      if (!h->wfold0T) EDGL_TRY(dev_alloc(h, &h->wfold0T, (size_t)2 * h->Ka * 4 * d));
      EDGL_TRY(launch_transpose(h->wfold0, h->Ka, 4 * d, h->wfold0T, st));
      EDGL_TRY(launch_tf32_lo(h->wfold0T, (long long)h->Ka * 4 * d, h->wfold0T + (size_t)h->Ka * 4 * d, st));
    }
  }
  // fp16 hi / lo / scale copies of the same K-major kernels for the scaled 3xFP16 dense layers (gemm_f16.cu)
  if (h->f16_gemm) {
    auto split16 = [&](std::map<std::string, unsigned char*>& dst, const std::string& key, const float* wt,
                       long long n) -> int {
      unsigned char*& buf = dst[key];
      if (!buf) EDGL_TRY(dev_alloc(h, &buf, w16_bytes(n)));
      return launch_w_split_f16(wt, n, buf, st);
    };
    for (int i = 0; i < h->cfg.num_blocks; ++i) {
      if (i > 0) EDGL_TRY(split16(h->bt16[i], "qkvt_w", h->btT[i].at("qkvt_w"), (long long)d * 4 * d));
      EDGL_TRY(split16(h->bt16[i], "ao_w", h->btT[i].at("ao_w"), (long long)d * d));
      EDGL_TRY(split16(h->bt16[i], "ff1_w", h->btT[i].at("ff1_w"), (long long)2 * d * d));
      EDGL_TRY(split16(h->bt16[i], "ff2_w", h->btT[i].at("ff2_w"), (long long)2 * d * d));
    }
    EDGL_TRY(split16(h->mt16, "tr_w", h->mtT.at("tr_w"), (long long)d * d));
    EDGL_TRY(split16(h->mt16, "wfold0", h->wfold0T, (long long)h->Ka * 4 * d));
    const long long nt = (h->c1 - h->c0) * (long long)d;
    if (nt * 4 <= (2ll << 30)) EDGL_TRY(split16(h->mt16, "table", F(h->mt, "item_embs") + h->c0 * d, nt));
  }
  int flag = 0;
  EDGL_CUDA(cudaMemcpyAsync(&flag, h->flag, sizeof(int), cudaMemcpyDeviceToHost, st));
  EDGL_CUDA(cudaStreamSynchronize(st));
  if (flag) return set_error(EDGL_EINVAL, "mark_table holds values outside [0, num_events=%d)", E);
  h->committed = true;
  return 0;
}

int edgl_encode(edgl_handle* h, const int64_t* seqs_i, const float* seqs_t, int B, float* y, void* stream) {
  EDGL_TRY(check_ready(h, B));
  if (!seqs_i || !seqs_t || !y) return set_error(EDGL_EINVAL, "null argument");
  return encode(h, seqs_i, seqs_t, B, y, (cudaStream_t)stream);
}

int edgl_encode_packed(edgl_handle* h, const int64_t* seqs_i, const float* seqs_t, int B, float* rows, int64_t row_stride,
                       void* stream) {
  EDGL_TRY(check_ready(h, B));
  if (!seqs_i || !seqs_t || !rows) return set_error(EDGL_EINVAL, "null argument");
  EDGL_REQUIRE(row_stride >= h->d + 2 * h->L && row_stride % 4 == 0, "row_stride must be a multiple of 4 >= d + 2 L");
  cudaStream_t st = (cudaStream_t)stream;
  EDGL_TRY(encode(h, seqs_i, seqs_t, B, rows, st, row_stride));
  // the ids ride behind y in the same row (raw bytes in fp32 lanes): one strided device-to-device copy
  EDGL_CUDA(cudaMemcpy2DAsync(rows + h->d, (size_t)row_stride * sizeof(float), seqs_i, (size_t)h->L * sizeof(int64_t),
                              (size_t)h->L * sizeof(int64_t), (size_t)B, cudaMemcpyDeviceToDevice, st));
  return 0;
}

int edgl_forward_logits(edgl_handle* h, const int64_t* seqs_i, const float* seqs_t, int B, float* logits,
                        void* stream) {
  EDGL_TRY(check_ready(h, B));
  if (!seqs_i || !seqs_t || !logits) return set_error(EDGL_EINVAL, "null argument");
  if (B == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  EDGL_TRY(encode(h, seqs_i, seqs_t, B, h->y, st));
  mark(h, ST_LOGITS_GEMM, st);
  EDGL_TRY(logits_rows(h, h->y, h->d, B, logits, (int)(h->c1 - h->c0), st));
  mark(h, ST_END, st);
  return 0;
}

static int check_train_args(edgl_handle* h, int B, const int64_t* positions, int M) {
  EDGL_TRY(check_ready(h, B));
  EDGL_REQUIRE(h->cfg.shard_world == 1, "the training-mode forward needs an unsharded handle");
  if (h->cfg.model == EDGL_MODEL_EASYDGL)
    EDGL_REQUIRE(positions && M >= 1 && M <= h->L, "EasyDGL training needs masked_positions [B, M], 1 <= M <= L");
  else
    EDGL_REQUIRE(!positions && M == h->L, "CTSMA predicts every position in training: positions = NULL, M = seqslen");
  return 0;
}

int edgl_forward_train_logits(edgl_handle* h, const int64_t* seqs_i, const float* seqs_t, int B,
                              const int64_t* masked_positions, int M, float* logits, void* stream) {
  EDGL_TRY(check_train_args(h, B, masked_positions, M));
  if (!seqs_i || !seqs_t || !logits) return set_error(EDGL_EINVAL, "null argument");
  if (B == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const float* Yg = nullptr;
  EDGL_TRY(train_encode(h, seqs_i, seqs_t, B, masked_positions, M, false, &Yg, st));
  mark(h, ST_LOGITS_GEMM, st);
  EDGL_TRY(logits_rows(h, Yg, h->d, (long long)B * M, logits, (int)(h->c1 - h->c0), st));
  mark(h, ST_END, st);
  int flag = 0;
  EDGL_CUDA(cudaMemcpyAsync(&flag, h->flag, sizeof(int), cudaMemcpyDeviceToHost, st));
  EDGL_CUDA(cudaStreamSynchronize(st));
  if (masked_positions && flag) return set_error(EDGL_EINVAL, "masked_positions holds values outside [0, L)");
  return 0;
}

int edgl_forward_train_loss(edgl_handle* h, const int64_t* seqs_i, const float* seqs_t, int B,
                            const int64_t* masked_positions, const int64_t* labels, int M, float l2_reg, float ct_reg,
                            float* loss_out, void* stream) {
  EDGL_TRY(check_train_args(h, B, masked_positions, M));
  if (!seqs_i || !seqs_t || !labels || !loss_out) return set_error(EDGL_EINVAL, "null argument");
  EDGL_REQUIRE(l2_reg >= 0.f, "Setting a scale less than 0 on a regularizer: %g.", (double)l2_reg);  // coding.py:27-29
  if (B == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const bool easy = h->cfg.model == EDGL_MODEL_EASYDGL;
  const int nb = h->cfg.num_blocks, d = h->d;
  const long long rows = (long long)B * M;
  const float* Yg = nullptr;
  EDGL_TRY(train_encode(h, seqs_i, seqs_t, B, masked_positions, M, ct_reg != 0.f, &Yg, st));
  // ---- masked softmax cross entropy over the item catalogue, the logits materialised in workspace-sized chunks
  const int Ns = (int)(h->c1 - h->c0);
  const int ldw = (Ns + 3) & ~3;
  if (!masked_positions) EDGL_CUDA(cudaMemsetAsync(h->flag, 0, sizeof(int), st));
  for (long long r0 = 0; r0 < rows; r0 += h->ws_rows) {
    const long long rc = (rows - r0 < h->ws_rows) ? (rows - r0) : h->ws_rows;
    mark(h, ST_LOGITS_GEMM, st);
    EDGL_TRY(logits_rows(h, Yg + r0 * d, d, rc, h->logits_ws, ldw, st));
    mark(h, ST_END, st);
    EDGL_TRY(launch_ce_rows(h->logits_ws, ldw, Ns, labels, r0, (int)rc, h->train_pe, h->train_wt, h->flag, st));
  }
  double* acc = h->train_acc;
  EDGL_TRY(launch_reduce(h->train_pe, rows, 0, 1.0, acc + 0, 0, st));
  EDGL_TRY(launch_reduce(h->train_wt, rows, 0, 1.0, acc + 1, 0, st));
  // ---- tf.losses.get_regularization_loss(): the embedding tables built with l2_reg (raw variables, row 0 included)
  EDGL_CUDA(cudaMemsetAsync(acc + 2, 0, sizeof(double), st));
  if (l2_reg != 0.f) {
    EDGL_TRY(launch_reduce(F(h->mt, "item_embs"), (long long)h->N1 * d, 1, 0.5 * l2_reg, acc + 2, 1, st));
    EDGL_TRY(launch_reduce(F(h->mt, "pos_embs"), (long long)h->L * d, 1, 0.5 * l2_reg, acc + 2, 1, st));
    if (easy) EDGL_TRY(launch_reduce(F(h->mt, "mark_embs"), (long long)h->E * d, 1, 0.5 * l2_reg, acc + 2, 1, st));
  }
  // ---- continuous-time regulariser: one biased likelihood per block (collection "LLE_PP")
  int nct = 0;
  if (ct_reg != 0.f) {
    const long long n = (long long)h->h * rows;
    float* ell = h->train_tpp;
    float* nu = ell + (size_t)h->h * h->train_rows_cap;
    float* cnt = nu + (size_t)h->h * h->train_rows_cap;
    for (int i = 0; i < nb; ++i) {
      EDGL_TRY(launch_tpp_terms(h->train_lam[i], masked_positions, labels, h->mark8, h->cfg.mark_rows, seqs_t, h->ts_len, B,
                                h->L, M, h->h, h->E, ell, nu, cnt, st));
      EDGL_TRY(launch_reduce(ell, n, 0, 1.0, acc + 3 + 3 * i + 0, 0, st));
      EDGL_TRY(launch_reduce(nu, n, 0, 1.0, acc + 3 + 3 * i + 1, 0, st));
      EDGL_TRY(launch_reduce(cnt, n, 0, 1.0, acc + 3 + 3 * i + 2, 0, st));
    }
    nct = nb;
  }
  // EasyDGL divides the regulariser by the number of heads (EasyDGL.py:175), CTSMA does not (CTSMA.py:110)
  EDGL_TRY(launch_loss_combine(acc, nct, (double)ct_reg / (easy ? (double)h->h : 1.0), loss_out, st));
  int flag = 0;
  EDGL_CUDA(cudaMemcpyAsync(&flag, h->flag, sizeof(int), cudaMemcpyDeviceToHost, st));
  EDGL_CUDA(cudaStreamSynchronize(st));
  if (flag == 1) return set_error(EDGL_EINVAL, "masked_positions holds values outside [0, L)");
  if (flag == 2) return set_error(EDGL_EINVAL, "labels hold values outside [0, num_items)");
  return 0;
}

int edgl_forward_topk(edgl_handle* h, const int64_t* seqs_i, const float* seqs_t, int B, int mask_seen,
                      int32_t* idx, float* val, void* stream) {
  EDGL_TRY(check_ready(h, B));
  if (!seqs_i || !seqs_t || !idx || !val) return set_error(EDGL_EINVAL, "null argument");
  EDGL_REQUIRE(h->cfg.shard_world == 1, "edgl_forward_topk needs an unsharded handle; use edgl_encode + "
               "edgl_logits_topk + edgl_topk_merge");
  if (B == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  EDGL_TRY(encode(h, seqs_i, seqs_t, B, h->y, st));
  return logits_topk(h, h->y, h->d, mask_seen ? seqs_i : nullptr, h->L, h->L, B, idx, val, 0, st);
}

int edgl_forward_topk_host_submit(edgl_handle* h, const int64_t* seqs_i_host, const float* seqs_t_host, int B,
                                  int mask_seen, int32_t* idx_host, float* val_host, void* stream) {
  EDGL_TRY(check_ready(h, B));
  if (!seqs_i_host || !seqs_t_host || !idx_host || !val_host) return set_error(EDGL_EINVAL, "null argument");
  const int sl = h->next_slot;
  EDGL_REQUIRE(!h->slot_busy[sl], "both staging slots are in flight: call edgl_forward_topk_host_wait first");
  if (B == 0) return sl;
  cudaStream_t st = (cudaStream_t)stream;
  // upload on cs_in (after the kernels that last read this slot's inputs), kernels on the caller's stream, download
  // on cs_out; pinned host buffers make all three asynchronous
  if (h->slot_used[sl]) EDGL_CUDA(cudaStreamWaitEvent(h->cs_in, h->ev_done[sl], 0));
  EDGL_CUDA(cudaMemcpyAsync(h->st_ids[sl], seqs_i_host, (size_t)B * h->L * sizeof(int64_t), cudaMemcpyHostToDevice, h->cs_in));
  EDGL_CUDA(cudaMemcpyAsync(h->st_ts[sl], seqs_t_host, (size_t)B * h->ts_len * sizeof(float), cudaMemcpyHostToDevice, h->cs_in));
  EDGL_CUDA(cudaEventRecord(h->ev_h2d[sl], h->cs_in));
  EDGL_CUDA(cudaStreamWaitEvent(st, h->ev_h2d[sl], 0));
  if (h->slot_used[sl]) EDGL_CUDA(cudaStreamWaitEvent(st, h->ev_d2h[sl], 0));  // the previous results have left the slot
  EDGL_TRY(edgl_forward_topk(h, h->st_ids[sl], h->st_ts[sl], B, mask_seen, h->st_idx[sl], h->st_val[sl], stream));
  EDGL_CUDA(cudaEventRecord(h->ev_done[sl], st));
  EDGL_CUDA(cudaStreamWaitEvent(h->cs_out, h->ev_done[sl], 0));
  EDGL_CUDA(cudaMemcpyAsync(idx_host, h->st_idx[sl], (size_t)B * h->K * sizeof(int32_t), cudaMemcpyDeviceToHost, h->cs_out));
  EDGL_CUDA(cudaMemcpyAsync(val_host, h->st_val[sl], (size_t)B * h->K * sizeof(float), cudaMemcpyDeviceToHost, h->cs_out));
  EDGL_CUDA(cudaEventRecord(h->ev_d2h[sl], h->cs_out));
  h->slot_busy[sl] = true;
  h->slot_used[sl] = true;
  h->next_slot = sl ^ 1;
  return sl;
}

int edgl_forward_topk_host_wait(edgl_handle* h, int slot) {
  if (!h) return set_error(EDGL_EINVAL, "null handle");
  EDGL_REQUIRE(slot == 0 || slot == 1, "slot must be 0 or 1");
  if (!h->slot_busy[slot]) return 0;
  EDGL_CUDA(cudaEventSynchronize(h->ev_d2h[slot]));
  h->slot_busy[slot] = false;
  return 0;
}

int edgl_forward_topk_host(edgl_handle* h, const int64_t* seqs_i_host, const float* seqs_t_host, int B,
                           int mask_seen, int32_t* idx_host, float* val_host, void* stream) {
  const int sl = edgl_forward_topk_host_submit(h, seqs_i_host, seqs_t_host, B, mask_seen, idx_host, val_host, stream);
  if (sl < 0) return sl;
  return edgl_forward_topk_host_wait(h, sl);
}

int edgl_logits_topk(edgl_handle* h, const float* y, int64_t y_stride, const int64_t* seen_ids, int seen_len,
                     int64_t seen_stride, int Bt, int64_t cand_stride, int32_t* cand_idx, float* cand_val,
                     void* stream) {
  EDGL_TRY(check_ready(h, Bt, false));
  if (!y || !cand_idx || !cand_val) return set_error(EDGL_EINVAL, "null argument");
  EDGL_REQUIRE(!seen_ids || seen_len >= 1, "seen_len must be >= 1");
  EDGL_REQUIRE(cand_stride == 0 || cand_stride >= h->K, "cand_stride must be 0 or >= K");
  EDGL_REQUIRE(y_stride == 0 || (y_stride >= h->d && y_stride % 4 == 0), "y_stride must be 0 or a multiple of 4 >= d");
  EDGL_REQUIRE(seen_stride == 0 || seen_stride >= seen_len, "seen_stride must be 0 or >= seen_len");
  return logits_topk(h, y, (int)y_stride, seen_ids, seen_len, seen_stride, Bt, cand_idx, cand_val, cand_stride,
                     (cudaStream_t)stream);
}

/* ---- fused exchange over peer memory (CUDA IPC + NVLink P2P stores) ---- */
int edgl_xchg_alloc(int64_t bytes, void** dev_ptr, void* ipc_handle_out) {
  if (!dev_ptr || !ipc_handle_out || bytes <= 0) return set_error(EDGL_EINVAL, "bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  void* p = nullptr;
  EDGL_CUDA(cudaMalloc(&p, (size_t)bytes));
  EDGL_CUDA(cudaMemset(p, 0, (size_t)bytes));
  cudaIpcMemHandle_t hd;
  EDGL_CUDA(cudaIpcGetMemHandle(&hd, p));
  memcpy(ipc_handle_out, &hd, sizeof(hd));
  *dev_ptr = p;
  return 0;
}

int edgl_xchg_open(const void* ipc_handle, void** dev_ptr) {
  if (!ipc_handle || !dev_ptr) return set_error(EDGL_EINVAL, "null argument");
  cudaIpcMemHandle_t hd;
  memcpy(&hd, ipc_handle, sizeof(hd));
  EDGL_CUDA(cudaIpcOpenMemHandle(dev_ptr, hd, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}

int edgl_xchg_close(void* dev_ptr) {
  if (dev_ptr) EDGL_CUDA(cudaIpcCloseMemHandle(dev_ptr));
  return 0;
}

int edgl_xchg_free(void* dev_ptr) {
  if (dev_ptr) EDGL_CUDA(cudaFree(dev_ptr));
  return 0;
}

int edgl_xchg_put_rows(edgl_handle* h, const float* y, int64_t y_stride, const int64_t* seqs_i, int B,
                       const int64_t* peer_rows, const int64_t* peer_flags, int G, int rank, uint32_t epoch,
                       void* stream) {
  EDGL_TRY(check_ready(h, B));
  if (!y || !seqs_i || !peer_rows || !peer_flags) return set_error(EDGL_EINVAL, "null argument");
  EDGL_REQUIRE(G >= 1 && G <= 32 && rank >= 0 && rank < G, "bad G/rank");
  if (y_stride == 0) y_stride = h->d;
  return launch_put_rows(y, y_stride, seqs_i, h->L, h->d, B, reinterpret_cast<const long long*>(peer_rows), G, rank,
                         reinterpret_cast<const long long*>(peer_flags), epoch, h->p2p_counter, (cudaStream_t)stream);
}

int edgl_xchg_wait(const uint32_t* flags, int G, uint32_t epoch, void* stream) {
  if (!flags) return set_error(EDGL_EINVAL, "null argument");
  EDGL_REQUIRE(G >= 1 && G <= 32, "bad G");
  return launch_wait_flags(flags, G, epoch, (cudaStream_t)stream);
}

int edgl_logits_topk_p2p(edgl_handle* h, const float* y, int64_t y_stride, const int64_t* seen_ids, int seen_len,
                         int64_t seen_stride, int Bt, int rows_per_dest, const int64_t* peer_cand,
                         const int64_t* peer_flags, int G, int rank, uint32_t epoch, void* stream) {
  EDGL_TRY(check_ready(h, Bt, false));
  if (!y || !peer_cand || !peer_flags) return set_error(EDGL_EINVAL, "null argument");
  EDGL_REQUIRE(G >= 1 && G <= 32 && rank >= 0 && rank < G && rows_per_dest >= 1 && Bt == G * rows_per_dest,
               "bad G/rank/rows_per_dest");
  EDGL_REQUIRE(y_stride == 0 || (y_stride >= h->d && y_stride % 4 == 0), "y_stride must be 0 or a multiple of 4 >= d");
  TopkP2P pp;
  memset(&pp, 0, sizeof(pp));
  pp.dest = reinterpret_cast<const long long*>(peer_cand);
  pp.peer_flags = reinterpret_cast<const long long*>(peer_flags);
  pp.counter = h->p2p_counter + 1;
  pp.rows_per_dest = rows_per_dest; pp.my_rank = rank; pp.G = G; pp.epoch = epoch;
  return logits_topk(h, y, (int)y_stride, seen_ids, seen_len, seen_stride, Bt, nullptr, nullptr, 0,
                     (cudaStream_t)stream, &pp);
}

int edgl_topk_merge(const float* cand_val, const int32_t* cand_idx, int G, int Bt, int K, int64_t shard_stride,
                    int64_t row_stride, int32_t* idx, float* val, void* stream) {
  if (!cand_val || !cand_idx || !idx || !val) return set_error(EDGL_EINVAL, "null argument");
  EDGL_REQUIRE(Bt >= 0 && shard_stride >= 0 && row_stride >= 0, "negative batch or stride");
  return launch_topk_merge(cand_val, cand_idx, G, Bt, K, shard_stride, row_stride, idx, val, (cudaStream_t)stream);
}

int edgl_time_sinusoid_code(const float* ts, int B, int L, int d, float* out, void* stream) {
  if (!ts || !out) return set_error(EDGL_EINVAL, "null argument");
  EDGL_REQUIRE(B >= 0 && L >= 0 && d >= 2 && d % 2 == 0, "time_sinusoid_code: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  std::vector<float> sc(d / 2);
  for (int j = 0; j < d / 2; ++j) sc[j] = (float)pow(10000.0, (double)(2 * j) * 1.0 / (double)d);
  float* dsc = nullptr;
  EDGL_CUDA(cudaMallocAsync(&dsc, sc.size() * sizeof(float), st));
  EDGL_CUDA(cudaMemcpyAsync(dsc, sc.data(), sc.size() * sizeof(float), cudaMemcpyHostToDevice, st));
  int rc = launch_time_code(ts, dsc, (long long)B * L, d, out, st);
  cudaFreeAsync(dsc, st);
  EDGL_CUDA(cudaStreamSynchronize(st));
  return rc;
}

int edgl_time_function_code(const float* x, const float* basis_freq, const float* phase, int64_t n, int d, float* out,
                            void* stream) {
  if (!x || !basis_freq || !phase || !out) return set_error(EDGL_EINVAL, "null argument");
  EDGL_REQUIRE(n >= 0 && d >= 1, "time_function_code: bad shape");
  return launch_time_function_code(x, basis_freq, phase, n, d, out, (cudaStream_t)stream);
}

int edgl_embedding_lookup(const float* table, int vocab, int d, int zero_pad, int scale, const int64_t* ids,
                          int64_t n_ids, float* out, void* stream) {
  if (!table || !ids || !out) return set_error(EDGL_EINVAL, "null argument");
  EDGL_REQUIRE(vocab >= 1 && d >= 1 && n_ids >= 0, "embedding_lookup: bad shape");
  const float s = scale ? (float)sqrt((double)d) : 1.0f;  // coding.py:62-63
  return launch_lookup(table, vocab, d, zero_pad, s, ids, n_ids, out, (cudaStream_t)stream);
}

int edgl_embed(edgl_handle* h, const int64_t* seqs_i, const float* seqs_t, int B, float* X0, float* spans,
               uint8_t* marks, void* stream) {
  EDGL_TRY(check_ready(h, B, false));
  if (!seqs_i || !seqs_t) return set_error(EDGL_EINVAL, "null argument");
  EmbedArgs e = embed_args(h, seqs_i, seqs_t, B);
  e.X0 = X0;
  e.ldx0 = (h->cfg.model == EDGL_MODEL_EASYDGL ? 3 : 2) * h->d;
  e.spans = spans;
  e.marks = marks;
  return launch_embed(e, (cudaStream_t)stream);
}

int edgl_attention_layer(edgl_handle* h, int block, const float* queries, int Cq, const float* keys, int Ck,
                         const uint8_t* kmask, const float* intervals, const uint8_t* marks, int B, int causality,
                         float* out, float* lam, void* stream) {
  EDGL_TRY(check_ready(h, B));
  if (!queries || !kmask || !intervals || !marks || !out) return set_error(EDGL_EINVAL, "null argument");
  EDGL_REQUIRE(block >= 0 && block < h->cfg.num_blocks, "block %d out of range", block);
  cudaStream_t st = (cudaStream_t)stream;
  const int d = h->d;
  const long long rows = (long long)B * h->L;
  const auto& w = h->bt[block];
  const int cin = cin_of(h, block);
  EDGL_REQUIRE(Cq == cin, "queries width %d does not match the block's dense kernel (%d)", Cq, cin);
  // projections: the tcgen05 GEMM the pipeline uses (default), or the exact-fp32 CUDA-core one (EDGL_LAYER_GEMM=simt)
  const char* lg = getenv("EDGL_LAYER_GEMM");
  const bool proj_tc = !(lg && lg[0] == 's');
  if (h->cfg.model == EDGL_MODEL_EASYDGL) {
    // BiMAU: `keys` and `causality` are ignored by the reference (temporal.py:404-429, Q15)
    if (proj_tc)
      EDGL_TRY(dense_nk(queries, Cq, h->btT[block].at(block > 0 ? "qkvt_w" : "qkvt_w_raw"), Cq, F(w, "qkvt_b"), h->qkvt,
                        4 * d, rows, 4 * d, ACT_NONE, nullptr, 0, st));
    else
      EDGL_TRY(dense(queries, Cq, F(w, "qkvt_w"), 4 * d, F(w, "qkvt_b"), h->qkvt, 4 * d, rows, 4 * d, Cq, ACT_NONE,
                     nullptr, 0, st));
    // bit 1 of `causality` selects T.MGAU (temporal.py:455-508): BiMAU without tf.linalg.set_diag
    AttnArgs a = attn_args(h, w, h->qkvt, kmask, intervals, marks, queries, Cq, out, lam, B, false,
                           (causality & 2) == 0);
    return launch_attention(a, st);
  }
  if (!keys) return set_error(EDGL_EINVAL, "MAU needs keys");
  EDGL_REQUIRE(Ck == cin, "keys width %d does not match the block's dense kernels (%d)", Ck, cin);
  if (proj_tc) {
    EDGL_TRY(dense_nk(queries, Cq, h->btT[block].at("q_w"), Cq, F(w, "q_b"), h->qkvt, 4 * d, rows, d, ACT_NONE, nullptr,
                      0, st));
    EDGL_TRY(dense_nk(keys, Ck, h->btT[block].at("kvt_w"), Ck, h->bkvt[block], h->qkvt + d, 4 * d, rows, 3 * d, ACT_NONE,
                      nullptr, 0, st));
  } else {
    EDGL_TRY(dense(queries, Cq, F(w, "q_w"), d, F(w, "q_b"), h->qkvt, 4 * d, rows, d, Cq, ACT_NONE, nullptr, 0, st));
    EDGL_TRY(dense(keys, Ck, h->wkvt[block], 3 * d, h->bkvt[block], h->qkvt + d, 4 * d, rows, 3 * d, Ck, ACT_NONE,
                   nullptr, 0, st));
  }
  AttnArgs a = attn_args(h, w, h->qkvt, kmask, intervals, marks, queries, Cq, out, lam, B, (causality & 1) != 0, false);
  return launch_attention(a, st);
}

int edgl_intensity(edgl_handle* h, int block, const float* H, const float* intervals, const uint8_t* marks, int B,
                   float* G, float* lam, void* stream) {
  EDGL_TRY(check_ready(h, B, false));
  if (!H || !intervals || !marks) return set_error(EDGL_EINVAL, "null argument");
  EDGL_REQUIRE(block >= 0 && block < h->cfg.num_blocks, "block %d out of range", block);
  const auto& w = h->bt[block];
  return launch_intensity(H, intervals, marks, F(w, "int_w"), F(w, "int_b"), F(w, "int_weight"), F(w, "int_scaling"),
                          B, h->L, h->h, h->dh, h->E, G, lam, (cudaStream_t)stream);
}

int edgl_layernorm(const float* x, const float* gamma, const float* beta, int B, int L, int C, float* out,
                   void* stream) {
  if (!x || !gamma || !beta || !out) return set_error(EDGL_EINVAL, "null argument");
  EDGL_REQUIRE(B >= 0 && L >= 1 && C >= 1, "layernorm: bad shape");
  return launch_layernorm(x, gamma, beta, B, L, C, out, false, (cudaStream_t)stream);
}

int edgl_time_attention(const float* Q, const float* K, const float* V, const uint8_t* key_mask,
                        const uint8_t* query_mask, const float* pos_k, const float* pos_v, int time_mode,
                        const void* intervals, const float* time_k, const float* time_v, int vocab,
                        const float* basis_freq, const float* phase, const float* U, float* TC, const float* residual,
                        int B, int Tq, int Tk, int C, int num_heads, int causality, float* out, void* stream) {
  if (!Q || !K || !V || !out) return set_error(EDGL_EINVAL, "null argument");
  EDGL_REQUIRE(B >= 0 && Tq >= 0 && Tk >= 1 && C >= 1, "time_attention: bad shape");
  return launch_time_attention(Q, K, V, key_mask, query_mask, pos_k, pos_v, time_mode, intervals, time_k, time_v, vocab,
                               basis_freq, phase, U, TC, residual, B, Tq, Tk, C, num_heads, causality, out,
                               (cudaStream_t)stream);
}

int edgl_row_nonzero(const float* x, int64_t rows, int C, uint8_t* out, void* stream) {
  if (!x || !out) return set_error(EDGL_EINVAL, "null argument");
  EDGL_REQUIRE(rows >= 0 && C >= 1, "row_nonzero: bad shape");
  return launch_row_nonzero(x, rows, C, out, (cudaStream_t)stream);
}

int edgl_layernorm_last(const float* x, const float* gamma, const float* beta, int64_t rows, int C, float eps, float* out,
                        void* stream) {
  if (!x || !gamma || !beta || !out) return set_error(EDGL_EINVAL, "null argument");
  EDGL_REQUIRE(rows >= 0 && C >= 1 && eps >= 0.f, "layernorm_last: bad shape");
  return launch_rownorm(x, gamma, beta, rows, C, eps, out, (cudaStream_t)stream);
}

int edgl_dense(const float* x, const float* w, const float* b, int M, int K, int N, int act, float* out,
               void* stream) {
  if (!x || !w || !out) return set_error(EDGL_EINVAL, "null argument");
  EDGL_REQUIRE(act >= 0 && act <= 2, "dense: unknown activation %d", act);
  return dense(x, K, w, N, b, out, N, M, N, K, act, nullptr, 0, (cudaStream_t)stream);
}

int edgl_dense_nk(const float* x, const float* wt, const float* b, int M, int K, int N, int act, float* out,
                  void* stream) {
  if (!x || !wt || !out) return set_error(EDGL_EINVAL, "null argument");
  EDGL_REQUIRE(act >= 0 && act <= 2, "dense: unknown activation %d", act);
  return dense_nk(x, K, wt, K, b, out, N, M, N, act, nullptr, 0, (cudaStream_t)stream, false);
}

int edgl_dense_nk_f16(const float* x, const float* wt, const float* b, int M, int K, int N, int act, float* out,
                      void* stream) {
  if (!x || !wt || !out) return set_error(EDGL_EINVAL, "null argument");
  EDGL_REQUIRE(act >= 0 && act <= 2, "dense: unknown activation %d", act);
  EDGL_REQUIRE(M >= 1 && N >= 1 && K >= 8 && K % 8 == 0, "dense_nk_f16: needs M, N >= 1 and K a multiple of 8");
  cudaStream_t st = (cudaStream_t)stream;
  // what edgl_commit (weight copies) and the producing kernel (activation maximum) do inside the pipeline
  unsigned char* buf = nullptr;
  const size_t wb = w16_bytes((long long)N * K);
  EDGL_CUDA(cudaMalloc(&buf, wb + 16));
  unsigned int* slot = reinterpret_cast<unsigned int*>(buf + wb);
  int rc = 0;
  if (cudaMemsetAsync(slot, 0, 16, st) != cudaSuccess) rc = set_error(EDGL_ECUDA, "cudaMemsetAsync failed");
  if (!rc) rc = launch_w_split_f16(wt, (long long)N * K, buf, st);
  if (!rc) rc = launch_absmax(x, (long long)M * K, slot, st);
  if (!rc) {
    GemmArgs g;
    g.A = x; g.lda = K; g.W = wt; g.ldw = K; g.w_is_nk = true; g.C = out; g.ldc = N;
    g.M = M; g.N = N; g.K = K; g.bias = b; g.act = act;
    g.W16 = buf; g.a_amax = slot;
    if (!gemm_f16_supported(g)) rc = set_error(EDGL_EINVAL, "dense_nk_f16: operands must be 16-byte aligned");
    else rc = launch_gemm_f16(g, st);
  }
  cudaStreamSynchronize(st);
  cudaFree(buf);
  return rc;
}

int edgl_topk(float* logits, int B, int N, const int64_t* seen_ids, int seen_len, int K, int32_t* idx, float* val,
              void* stream) {
  if (!logits || !idx || !val) return set_error(EDGL_EINVAL, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (seen_ids) EDGL_TRY(launch_mask_seen(logits, N, B, seen_ids, seen_len, seen_len, 0, N, st));
  return launch_topk(logits, N, B, N, K, 0, 0, idx, val, st);
}

}  // extern "C"
