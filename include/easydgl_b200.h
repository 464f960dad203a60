/* easydgl_b200.h - C ABI of libeasydgl_b200.so (sm_100a).
 *
 * The reference (cchao0116/EasyDGL) has no FFI / plugin boundary: the boundary for
 * the hot path is the Python object protocol of its model and layer classes
 * (SURVEY.md section 8b).  Each entry point below names the reference interface it
 * replaces (file:line relative to the reference tree).  The Python facade in
 * easydgl_b200/{model,module}/ keeps the reference's class names and call
 * signatures and binds these symbols with ctypes (INTEGRATION.md shows the stub).
 *
 * Conventions
 *  - every pointer named *_dev / without suffix is a DEVICE pointer unless the
 *    function name ends in _host; tensors are dense row-major fp32 unless stated;
 *    ids are int64 (the reference feeds tf.int64 seqs_i, dataloader.py:14-26).
 *  - every call is asynchronous on the given stream (a cudaStream_t passed as
 *    void*; NULL = the legacy default stream) except the *_host calls, which
 *    synchronise that stream before returning.
 *  - return value: 0 = ok, negative = error (EDGL_E*); edgl_last_error() returns a
 *    thread-local message.  No call allocates device memory except edgl_create,
 *    edgl_commit and edgl_reserve.
 *  - weights are BORROWED: the caller keeps the device buffers alive for the life
 *    of the handle (or until the next edgl_set_tensor for that name).
 */
#ifndef EASYDGL_B200_H
#define EASYDGL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EDGL_OK 0
#define EDGL_EINVAL (-1)   /* bad argument / shape / alignment */
#define EDGL_ESTATE (-2)   /* call order (e.g. forward before commit, missing tensor) */
#define EDGL_ECUDA (-3)    /* CUDA runtime error (message holds cudaGetErrorString) */
#define EDGL_ENOMEM (-4)   /* workspace allocation failed */
#define EDGL_EARCH (-5)    /* device is not sm_100 */

#define EDGL_MODEL_EASYDGL 0 /* src/model/EasyDGL.py */
#define EDGL_MODEL_CTSMA 1   /* src/model/CTSMA.py */

typedef struct edgl_handle edgl_handle;

/* Mirrors what Model(num_items, FLAGS) derives in its ctor
 * (EasyDGL.py:37-67, CTSMA.py:22-44, Base.py:92-104). */
typedef struct edgl_config {
  int32_t model;       /* EDGL_MODEL_* */
  int32_t max_batch;   /* workspace is sized for this many sequences per call */
  int32_t seq_len;     /* L: length of seqs_i (EasyDGL: FLAGS.seqslen+1, EasyDGL.py:40; CTSMA: FLAGS.seqslen) */
  int32_t num_units;   /* d  (FLAGS.num_units) */
  int32_t num_heads;   /* h  (FLAGS.num_heads) */
  int32_t num_blocks;  /* FLAGS.num_blocks */
  int32_t num_events;  /* E = mark_lookup_table.shape[-1] (EasyDGL.py:46) */
  int32_t num_rows;    /* rows of the item table = logit columns (EasyDGL: num_items+1, EasyDGL.py:41) */
  int32_t mark_rows;   /* rows of mark_lookup_table */
  int32_t topk;        /* K of tf.nn.top_k (Base.py:181 hard-codes 100) */
  float time_scale;    /* FLAGS.time_scale (EasyDGL.py:43) */
  int64_t mask_id;     /* [MASK] token id = FLAGS.num_items for EasyDGL (EasyDGL.py:39); -1 for CTSMA */
  int32_t shard_rank;  /* this handle owns logit columns [rank*ceil(N1/world), ...) - SURVEY 8e */
  int32_t shard_world; /* 1 = unsharded */
} edgl_config;

const char* edgl_last_error(void);
int edgl_version(void);
/* Number of CUDA kernels this library has launched in this process (bench.py's gpu_launches). */
int64_t edgl_launch_count(void);
/* Host-side CRC32C (Castagnoli) of data[0..n), continuing from `crc` (0 to start).  Replaces the checksum
 * TensorFlow applies to TFRecord frames (tf.data.TFRecordDataset, dataloader.py:230-236) and to the
 * tensor bundles of tf.train.Saver (util.py:26,53-55); no device work, callable without a GPU. */
uint32_t edgl_crc32c(const void* data, size_t n, uint32_t crc);

/* Optional per-stage device timing (no reference counterpart; the reference has no profiler).
 * While enabled, a CUDA event is recorded on the caller's stream before every kernel of the pipeline;
 * edgl_profile_read synchronises on the last event and returns, per stage, the summed device time (ms)
 * and the launch count since the last read.  Stage ids are 0..edgl_num_stages()-1. */
int edgl_num_stages(void);
const char* edgl_stage_name(int stage);
int edgl_profile(edgl_handle* h, int enable);
int edgl_profile_read(edgl_handle* h, double* ms, int64_t* count, int n);

/* Model(num_items, FLAGS) - EasyDGL.py:37 / CTSMA.py:22.  Uses the current CUDA device. */
int edgl_create(const edgl_config* cfg, edgl_handle** out);
int edgl_destroy(edgl_handle* h);
int edgl_get_config(const edgl_handle* h, edgl_config* out);

/* Bind one variable of the reference graph (tf.get_variable / tf.layers.dense kernels)
 * by name; block = -1 for model-level tensors.  numel is checked against the config.
 * Names:
 *   model level : item_embs [N1,d]  pos_embs [L,d]  output_bias [N1-1]  mark_table int64 [mark_rows,E]
 *     EasyDGL   : mark_embs [E,d]  tr_w [d,d] tr_b tr_ln_g tr_ln_b [d]
 *     CTSMA     : out_ln_g out_ln_b [d]
 *   per block   : int_w [dh+1,dh*E] int_b [dh*E] int_weight [E,dh] int_scaling [E]
 *     EasyDGL   : qkvt_w [Cin,4d] qkvt_b [4d]  ao_w [d,d] ao_b ao_ln_g ao_ln_b
 *                 ff1_w [d,2d] ff1_b  ff2_w [2d,d] ff2_b  ff_ln_g ff_ln_b      (Cin = 3d for block 0 else d)
 *     CTSMA     : ln1_g ln1_b [Cin]  {q,k,v,t}_w [Cin,d] {q,k,v,t}_b [d]  ln2_g ln2_b [d]
 *                 ff1_w [d,d] ff1_b ff2_w [d,d] ff2_b                          (Cin = 2d for block 0 else d)
 * All fp32 except mark_table (int64, values in [0,E): they index mark_embs, EasyDGL.py:87). */
int edgl_set_tensor(edgl_handle* h, const char* name, int block, const void* dev_ptr, int64_t numel);
/* Derive the device-side constants that depend only on weights (packed K|V|T kernels,
 * position/mark contributions of the block-0 QKVT dense, uint8 mark table).  Call after
 * the last edgl_set_tensor and again whenever a bound tensor's contents change. */
int edgl_commit(edgl_handle* h, void* stream);

/* model(features, is_training=False) -> logits [B, N1]        (EasyDGL.py:69-151 / CTSMA.py:46-91)
 * seqs_i int64 [B,L]; seqs_t fp32 [B,L] (CTSMA: [B,L+1], dataloader.py:99). */
int edgl_forward_logits(edgl_handle* h, const int64_t* seqs_i, const float* seqs_t, int B, float* logits,
                        void* stream);
/* model.eval(features, labels, mask_seen) ranking part (Base.py:150-181): forward, optional
 * -inf at every id in seqs_i, top-K with ties -> lower index.  idx int32 [B,K], val fp32 [B,K]
 * (the masked logits; ranking on them equals ranking on softmax).  Unsharded handles only. */
int edgl_forward_topk(edgl_handle* h, const int64_t* seqs_i, const float* seqs_t, int B, int mask_seen,
                      int32_t* idx, float* val, void* stream);
/* Same call with HOST buffers (pinned or pageable): H2D of seqs_i/seqs_t, forward, D2H of idx/val,
 * stream synchronised on return.  This is the end-to-end call bench.py times as "e2e". */
int edgl_forward_topk_host(edgl_handle* h, const int64_t* seqs_i_host, const float* seqs_t_host, int B,
                           int mask_seen, int32_t* idx_host, float* val_host, void* stream);
/* The same call split in two, so that a caller that feeds batch after batch (the reference's eval loop,
 * main.py:133-143, one sess.run per batch) keeps TWO batches in flight: submit enqueues upload (side stream),
 * kernels (`stream`) and download (second side stream) of one batch and returns its staging slot (0 or 1, or a
 * negative error code) without waiting; wait blocks until that slot's idx/val have arrived in the host buffers.
 * With pinned host buffers the upload of batch i+1 and the download of batch i-1 overlap the kernels of batch i.
 * At most two submits may be outstanding. */
int edgl_forward_topk_host_submit(edgl_handle* h, const int64_t* seqs_i_host, const float* seqs_t_host, int B,
                                  int mask_seen, int32_t* idx_host, float* val_host, void* stream);
int edgl_forward_topk_host_wait(edgl_handle* h, int slot);

/* ---- training-mode forward: model(features, is_training=True) and the loss of model.train() ----
 * Both dropout rates must be 0 (a TF-identical random stream cannot be reproduced; the backward pass and the
 * optimizer step are outside this library).  EasyDGL: masked_positions int64 [B,M] (features['masked_positions'],
 * dataloader.py:181-201), labels int64 [B,M]; CTSMA: masked_positions = NULL, M = seqslen, labels int64 [B,S]
 * (dataloader.py:95-98).  Unsharded handles only. */
/* logits [B*M, N1] at the predicted positions (EasyDGL.py:140-151 / CTSMA.py:82-91). */
int edgl_forward_train_logits(edgl_handle* h, const int64_t* seqs_i, const float* seqs_t, int B,
                              const int64_t* masked_positions, int M, float* logits, void* stream);
/* The scalar loss of model.train (EasyDGL.py:153-189 / CTSMA.py:93-124): masked softmax cross entropy
 * -log(softmax + 1e-5) over labels != 0, plus l2_reg * l2_loss of the embedding tables (coding.py:13-44), plus
 * ct_reg * MAU.biased_likelihood (temporal.py:317-333) of every block's intensities.  loss_out: DEVICE float[4] =
 * {loss, cross entropy, l2 regulariser, continuous-time regulariser}.  Synchronises the stream (it reports
 * out-of-range positions / labels). */
int edgl_forward_train_loss(edgl_handle* h, const int64_t* seqs_i, const float* seqs_t, int B,
                            const int64_t* masked_positions, const int64_t* labels, int M, float l2_reg, float ct_reg,
                            float* loss_out, void* stream);

/* ---- the two halves of the forward, for the column-sharded multi-GPU path ---- */
/* Encoder up to y = hidden[:, -1]  [B,d]  (EasyDGL.py:69-146 / CTSMA.py:46-87). */
int edgl_encode(edgl_handle* h, const int64_t* seqs_i, const float* seqs_t, int B, float* y, void* stream);
/* The encoder writing straight into the packed exchange rows of the multi-GPU path: row b of `rows` (row_stride
 * floats apart, >= d + 2 L) receives [y (d fp32) | seqs_i[b] (L int64, raw bytes)] - the message every rank
 * all-gathers, with no packing pass in between. */
int edgl_encode_packed(edgl_handle* h, const int64_t* seqs_i, const float* seqs_t, int B, float* rows, int64_t row_stride,
                       void* stream);
/* y [Bt,d] x this handle's item-table shard -> logits + bias (EasyDGL.py:149-150), optional seen-mask
 * with seen_ids int64 [Bt,seen_len] (may be NULL), local top-K with GLOBAL column ids.
 * Strides are in elements, 0 = dense: row r of y starts at y[r*y_stride] (multiple of 4), of seen_ids at
 * seen_ids[r*seen_stride], and its candidates are written at cand_*[r*cand_stride] - so the packed
 * [y | ids] exchange rows and an interleaved [idx | val] buffer are used in place, without copies. */
int edgl_logits_topk(edgl_handle* h, const float* y, int64_t y_stride, const int64_t* seen_ids, int seen_len,
                     int64_t seen_stride, int Bt, int64_t cand_stride, int32_t* cand_idx, float* cand_val,
                     void* stream);
/* K-way merge of G per-shard candidate lists -> [Bt,K]; ties -> lower global index (Base.py:181).
 * Shard g's block starts at cand_*[g * shard_stride], its rows are row_stride apart (elements; 0 means
 * dense [G,Bt,K]); idx < 0 = padding. */
int edgl_topk_merge(const float* cand_val, const int32_t* cand_idx, int G, int Bt, int K, int64_t shard_stride,
                    int64_t row_stride, int32_t* idx, float* val, void* stream);

/* ---- fused exchange over peer memory (no reference counterpart: the reference is single-GPU) ----
 * One process per GPU.  Each rank allocates an exchange region with edgl_xchg_alloc (cudaMalloc + CUDA IPC
 * handle, zero-filled), ships the 64-byte handle to its peers (any transport) and maps theirs with
 * edgl_xchg_open.  Kernels then write straight into peer memory over NVLink and raise per-source uint32
 * flags (system-scope release / acquire); there is no NCCL call on the data path.
 *   rows region   float  [G*B][d + 2L]  : slot r*B.. holds rank r's packed [y | seqs_i] rows
 *   cand region   int32  [G][B][2][K]   : block g holds shard g's candidates (idx | val bits) for MY rows
 *   flags         uint32 [G] per region : flag[g] = last epoch fully written by rank g */
int edgl_xchg_alloc(int64_t bytes, void** dev_ptr, void* ipc_handle_out /* 64 bytes */);
int edgl_xchg_open(const void* ipc_handle /* 64 bytes */, void** dev_ptr);
int edgl_xchg_close(void* dev_ptr);
int edgl_xchg_free(void* dev_ptr);
/* Write this rank's B packed rows into slot `rank` of every peer's rows region (peer_rows: device array of G
 * addresses) and then publish `epoch` in slot `rank` of every peer's flag array (peer_flags: device array). */
int edgl_xchg_put_rows(edgl_handle* h, const float* y, int64_t y_stride, const int64_t* seqs_i, int B,
                       const int64_t* peer_rows, const int64_t* peer_flags, int G, int rank, uint32_t epoch,
                       void* stream);
/* Enqueue a wait until all G local flags have reached `epoch`. */
int edgl_xchg_wait(const uint32_t* flags, int G, uint32_t epoch, void* stream);
/* edgl_logits_topk whose top-K kernel writes the candidates of row R directly into the cand region of rank
 * R / rows_per_dest (block `rank`) and publishes `epoch` in the peers' cand flags when the step is complete. */
int edgl_logits_topk_p2p(edgl_handle* h, const float* y, int64_t y_stride, const int64_t* seen_ids, int seen_len,
                         int64_t seen_stride, int Bt, int rows_per_dest, const int64_t* peer_cand,
                         const int64_t* peer_flags, int G, int rank, uint32_t epoch, void* stream);

/* ---- layer-level entry points (one per reference layer, for unit parity) ---- */
/* C.TimeSinusoidCoding(d).code(ts)  (coding.py:132-149): ts fp32 [B,L] already scaled -> [B,L,d]. */
int edgl_time_sinusoid_code(const float* ts, int B, int L, int d, float* out, void* stream);
/* C.TimeFunctionCoding(d).code(x)  (coding.py:97-122, the Bochner/Mercer harmonic kernel of TGAT; SURVEY 8f
 * rank 4): x fp32 [n] (any leading shape, flattened) -> out [n,d] = cos(x * basis_freq + phase). */
int edgl_time_function_code(const float* x, const float* basis_freq, const float* phase, int64_t n, int d, float* out,
                            void* stream);
/* Attention core of T.TiMultiHeadAttention (temporal.py:15-109; time_mode 1), T.TfMultiHeadAttention
 * (temporal.py:112-185; time_mode 2) and T.TgMultiHeadAttention (temporal.py:188-264; time_mode 3) after their dense
 * projections; time_mode 0 is plain multi-head attention.  Q [B,Tq,C], K / V [B,Tk,C] fp32; key_mask [B,Tk] /
 * query_mask [B,Tq] uint8 or null (temporal.py:65-70, 87-90); pos_k / pos_v [Tk,C] position codes added to K / V or null
 * (temporal.py:49-50, 57, 98); causality != 0 = future blinding (temporal.py:73-79).
 *   mode 1: intervals int64 [B,Tq,Tk] (clipped to [0, vocab)), time_k / time_v [vocab,C] interval embeddings
 *           (time_v may be null): s += Q . time_k[iv], o += P time_v[iv]            (temporal.py:51-52, 58, 99)
 *   mode 2: intervals fp32 [B,Tq,Tk]: s += Q . cos(iv basis_freq + phase) (head slice)  (temporal.py:141, 146)
 *   mode 3: intervals fp32, U [B,Tq,h,C] = the time half of the key projection applied to the query (W_k2[:, head]
 *           Q[head]): s += U . cos(iv basis_freq + phase); TC [B,Tq,h,C] (output) = sum_k P cos(.), the time half of
 *           the value product before W_v2                                         (temporal.py:212-220, 249-251)
 * out [B,Tq,C] = softmax-weighted values + residual (null = none).  Exact fp32. */
int edgl_time_attention(const float* Q, const float* K, const float* V, const uint8_t* key_mask,
                        const uint8_t* query_mask, const float* pos_k, const float* pos_v, int time_mode,
                        const void* intervals, const float* time_k, const float* time_v, int vocab,
                        const float* basis_freq, const float* phase, const float* U, float* TC, const float* residual,
                        int B, int Tq, int Tk, int C, int num_heads, int causality, float* out, void* stream);
/* tf.sign(tf.reduce_sum(tf.abs(x), -1)) of the key / query masking (temporal.py:65, 87): x [rows,C] -> out [rows] 0/1. */
int edgl_row_nonzero(const float* x, int64_t rows, int C, uint8_t* out, void* stream);
/* module.normalize.layernorm (normalize.py:9-19): moments over the last axis,
 * gamma (x - mean) / sqrt(var + eps) + beta; x [rows,C]. */
int edgl_layernorm_last(const float* x, const float* gamma, const float* beta, int64_t rows, int C, float eps, float* out,
                        void* stream);
/* C.Embedding(vocab,d,zero_pad,scale)(ids)  (coding.py:45-64): table [vocab,d] raw variable. */
int edgl_embedding_lookup(const float* table, int vocab, int d, int zero_pad, int scale, const int64_t* ids,
                          int64_t n_ids, float* out, void* stream);
/* EasyDGL.__call__ input assembly (EasyDGL.py:70-95) / CTSMA (CTSMA.py:47-60):
 * X0 [B,L,3d] (CTSMA [B,L,2d]); spans fp32 [B,L]; marks uint8 [B,L,E]. Any output may be NULL. */
int edgl_embed(edgl_handle* h, const int64_t* seqs_i, const float* seqs_t, int B, float* X0, float* spans,
               uint8_t* marks, void* stream);
/* T.BiMAU(...)(queries, keys, masks, intervals, marks, is_training=False)  (temporal.py:404-452)
 * and T.MAU(...)(..., causality)  (temporal.py:335-390), using block `block`'s weights.
 * queries [B,L,Cq]; keys [B,L,Ck] (ignored for EasyDGL handles, Q15); kmask uint8 [B,L] (1 = real key);
 * intervals [B,L]; marks uint8 [B,L,E].  Outputs: out [B,L,d]; lam [h*B,L,E] head-major (may be NULL).
 * causality: bit 0 = causal mask (MAU only); bit 1 = no set_diag on an EasyDGL handle = T.MGAU
 * (temporal.py:455-508, identical to BiMAU otherwise). */
int edgl_attention_layer(edgl_handle* h, int block, const float* queries, int Cq, const float* keys, int Ck,
                         const uint8_t* kmask, const float* intervals, const uint8_t* marks, int B,
                         int causality, float* out, float* lam, void* stream);
/* T.MAU.intensity(H, intervals, mark_onehot)  (temporal.py:281-315) with block `block`'s weights:
 * H [h*B,L,dh] head-major -> G [h*B,L,L] (no set_diag), lam [h*B,L,E]. */
int edgl_intensity(edgl_handle* h, int block, const float* H, const float* intervals, const uint8_t* marks,
                   int B, float* G, float* lam, void* stream);
/* Base.layernorm(x)  (Base.py:12-67): joint (L,C) statistics per sample. x [B,L,C] -> out [B,L,C]. */
int edgl_layernorm(const float* x, const float* gamma, const float* beta, int B, int L, int C, float* out,
                   void* stream);
/* tf.layers.dense(x, N, activation) (act: 0 none, 1 gelu-erf EasyDGL.py:19-32, 2 relu): x [M,K] @ w [K,N] + b. */
int edgl_dense(const float* x, const float* w, const float* b, int M, int K, int N, int act, float* out,
               void* stream);
/* Same layer with the kernel stored K-major, wt [N,K] = w^T (how edgl_commit keeps every dense kernel):
 * the tcgen05 tensor-core path (3xTF32, fp32-level accuracy). */
int edgl_dense_nk(const float* x, const float* wt, const float* b, int M, int K, int N, int act, float* out,
                  void* stream);
/* Same layer on the scaled 3xFP16 tensor-core path (tcgen05 kind::f16; fp32-level accuracy relative to
 * max|x| * |w|): K must be a multiple of 8.  Inside the model pipeline the weight copies are made by edgl_commit and
 * the activation maximum comes from the kernel that produced x; this stand-alone entry computes both first and
 * synchronises the stream before returning (a test / tooling entry, not a hot-path one). */
int edgl_dense_nk_f16(const float* x, const float* wt, const float* b, int M, int K, int N, int act, float* out,
                      void* stream);
/* Sequential.eval ranking on given logits (Base.py:156-181): logits [B,N] are modified in place
 * (seen ids -> -inf) when seen_ids != NULL. */
int edgl_topk(float* logits, int B, int N, const int64_t* seen_ids, int seen_len, int K, int32_t* idx,
              float* val, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EASYDGL_B200_H */
