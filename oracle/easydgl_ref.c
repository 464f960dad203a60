/* easydgl_ref.c - independent plain-C (double precision) restatement of the EasyDGL / CTSMA eval
 * forward pass.  TEST INFRASTRUCTURE ONLY: it cross-checks oracle/easydgl_oracle.py (two restatements
 * written separately from the reference's source text agreeing is the strongest pin available -
 * PARITY UNPINNED: the reference has no tests/golden vectors and TensorFlow 2.3.4 cannot run here).
 * Never linked into or called by the product.
 *
 * Follows, line by line (paths relative to /root/reference):
 *   src/model/EasyDGL.py:69-151, src/model/CTSMA.py:46-91, src/module/temporal.py:281-315,335-452,
 *   src/module/coding.py:45-79,125-149, src/model/Base.py:12-67,70-87,106-113.
 *
 * usage: easydgl_ref in.bin out.bin     (layout written by tests/test_c_oracle.py)
 *   header: int32 model(0 EasyDGL,1 CTSMA) B L d h blocks E N1 mark_rows ts_len; float64 time_scale; int64 mask_id
 *   then the arrays in the order read below (float32 unless noted); output: float64 logits [B,N1].
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static FILE* fin;
static double* rd(size_t n) { /* float32 array -> double */
  float* t = (float*)malloc(n * sizeof(float));
  double* o = (double*)malloc((n ? n : 1) * sizeof(double));
  if (fread(t, sizeof(float), n, fin) != n) { fprintf(stderr, "short read\n"); exit(2); }
  for (size_t i = 0; i < n; ++i) o[i] = t[i];
  free(t);
  return o;
}
static int64_t* rdi(size_t n) {
  int64_t* o = (int64_t*)malloc((n ? n : 1) * sizeof(int64_t));
  if (fread(o, sizeof(int64_t), n, fin) != n) { fprintf(stderr, "short read\n"); exit(2); }
  return o;
}

/* tf.layers.dense: out[L,co] = x[L,ci(ldx)] @ W[ci,co] + b */
static void dense(const double* x, int ldx, int L, int ci, const double* W, const double* b, int co, double* out) {
  for (int l = 0; l < L; ++l)
    for (int o = 0; o < co; ++o) {
      double s = b ? b[o] : 0.0;
      for (int i = 0; i < ci; ++i) s += x[(size_t)l * ldx + i] * W[(size_t)i * co + o];
      out[(size_t)l * co + o] = s;
    }
}
/* Base.layernorm (Base.py:12-67): moments over (L,C) jointly, population variance, eps 1e-12 */
static void layernorm(const double* x, int L, int C, const double* g, const double* be, double* out) {
  const size_t n = (size_t)L * C;
  double mean = 0, var = 0;
  for (size_t i = 0; i < n; ++i) mean += x[i];
  mean /= (double)n;
  for (size_t i = 0; i < n; ++i) var += (x[i] - mean) * (x[i] - mean);
  var /= (double)n;
  const double r = 1.0 / sqrt(var + 1e-12);
  for (int l = 0; l < L; ++l)
    for (int c = 0; c < C; ++c) {
      const double inv = r * g[c];
      out[(size_t)l * C + c] = x[(size_t)l * C + c] * inv + (be[c] - mean * inv); /* Base.py:57-63 */
    }
}
static double gelu(double x) { return x * 0.5 * (1.0 + erf(x / sqrt(2.0))); } /* EasyDGL.py:31-32 */

typedef struct { const double *w, *b, *weight, *scaling; } Intensity;

/* attention core of MAU/BiMAU for ONE sequence (temporal.py:345-382 / 412-444).
 * Q,K,V,T [L,d]; kmask[L]; spans[L]; marks[L,E]; out[L,d] (no residual). */
static void attention(const double* Q, const double* K, const double* V, const double* T, const double* kmask,
                      const double* spans, const int64_t* marks, Intensity iw, int L, int d, int h, int E, int causal,
                      int diag_one, double* out) {
  const int dh = d / h;
  const double fill = (double)(-4294967296.0 + 1.0); /* -2**32+1 */
  double* P = (double*)malloc((size_t)L * L * sizeof(double));
  double* H = (double*)malloc((size_t)dh * sizeof(double));
  double* lam = (double*)malloc((size_t)L * E * sizeof(double));
  for (int hd = 0; hd < h; ++hd) {
    const int o = hd * dh; /* tf.split(., h, axis=2)[hd] */
    for (int q = 0; q < L; ++q) {
      double m = -INFINITY, sum = 0;
      for (int k = 0; k < L; ++k) {
        double s = 0;
        for (int j = 0; j < dh; ++j) s += Q[(size_t)q * d + o + j] * K[(size_t)k * d + o + j];
        s /= sqrt((double)dh);                       /* :355 / :422 */
        if (kmask[k] == 0.0) s = fill;               /* :358-359 / :425-426 */
        if (causal && k > q) s = fill;               /* :362-367 */
        P[(size_t)q * L + k] = s;
        if (s > m) m = s;
      }
      for (int k = 0; k < L; ++k) { P[(size_t)q * L + k] = exp(P[(size_t)q * L + k] - m); sum += P[(size_t)q * L + k]; }
      for (int k = 0; k < L; ++k) P[(size_t)q * L + k] /= sum; /* softmax :370 / :429 */
      /* H = P T_ (:375 / :434) then MAU.intensity (:281-307) */
      for (int j = 0; j < dh; ++j) {
        double s = 0;
        for (int k = 0; k < L; ++k) s += P[(size_t)q * L + k] * T[(size_t)k * d + o + j];
        H[j] = s;
      }
      for (int e = 0; e < E; ++e) {
        double acc = 0;
        for (int j = 0; j < dh; ++j) {
          const int c = e * dh + j; /* tf.split(., E, axis=2)[e][:, :, j] */
          double z = iw.b[c] + spans[q] * iw.w[(size_t)dh * dh * E + c];
          for (int i = 0; i < dh; ++i) z += H[i] * iw.w[(size_t)i * dh * E + c];
          acc += (1.0 / (1.0 + exp(-z))) * iw.weight[(size_t)e * dh + j];
        }
        const double s = exp(iw.scaling[e]);
        lam[(size_t)q * E + e] = s * log(1.0 + exp(acc / s)); /* :305-306 */
      }
    }
    for (int q = 0; q < L; ++q)
      for (int j = 0; j < dh; ++j) {
        double s = 0;
        for (int k = 0; k < L; ++k) {
          double g = 0; /* :309-313 */
          for (int e = 0; e < E; ++e) g += lam[(size_t)q * E + e] * (double)marks[(size_t)k * E + e];
          if (diag_one && k == q) g = 1.0; /* :438-439 */
          s += g * P[(size_t)q * L + k] * V[(size_t)k * d + o + j]; /* :441-443 */
        }
        out[(size_t)q * d + o + j] = s;
      }
  }
  free(P); free(H); free(lam);
}

int main(int argc, char** argv) {
  if (argc != 3) { fprintf(stderr, "usage: %s in.bin out.bin\n", argv[0]); return 2; }
  fin = fopen(argv[1], "rb");
  if (!fin) { perror("in"); return 2; }
  int32_t hd[10];
  double time_scale; int64_t mask_id;
  if (fread(hd, 4, 10, fin) != 10 || fread(&time_scale, 8, 1, fin) != 1 || fread(&mask_id, 8, 1, fin) != 1) return 2;
  const int model = hd[0], B = hd[1], L = hd[2], d = hd[3], h = hd[4], nb = hd[5], E = hd[6], N1 = hd[7],
            mark_rows = hd[8], ts_len = hd[9], dh = d / h;
  int64_t* ids = rdi((size_t)B * L);
  double* ts = rd((size_t)B * ts_len);
  double* item = rd((size_t)N1 * d);
  double* pos = rd((size_t)L * d);
  double* obias = rd((size_t)N1 - 1);
  int64_t* mtab = rdi((size_t)mark_rows * E);
  double* memb = model == 0 ? rd((size_t)E * d) : NULL;
  for (int j = 0; j < d; ++j) item[j] = 0.0;               /* zero_pad (coding.py:56-57) */
  if (memb) for (int j = 0; j < d; ++j) memb[j] = 0.0;
  const int W0 = model == 0 ? 3 * d : 2 * d;
  typedef struct {
    double *qkvt_w, *qkvt_b, *q_w, *q_b, *k_w, *k_b, *v_w, *v_b, *t_w, *t_b, *ln1_g, *ln1_b, *ln2_g, *ln2_b, *ao_w, *ao_b,
        *ao_g, *ao_be, *ff1_w, *ff1_b, *ff2_w, *ff2_b, *ff_g, *ff_be;
    Intensity iw;
  } Blk;
  Blk* blk = (Blk*)calloc(nb, sizeof(Blk));
  for (int i = 0; i < nb; ++i) {
    const int cin = i == 0 ? W0 : d;
    Blk* k = &blk[i];
    if (model == 0) {
      k->qkvt_w = rd((size_t)cin * 4 * d); k->qkvt_b = rd(4 * d);
    } else {
      k->ln1_g = rd(cin); k->ln1_b = rd(cin);
      k->q_w = rd((size_t)cin * d); k->q_b = rd(d); k->k_w = rd((size_t)cin * d); k->k_b = rd(d);
      k->v_w = rd((size_t)cin * d); k->v_b = rd(d); k->t_w = rd((size_t)cin * d); k->t_b = rd(d);
    }
    k->iw.w = rd((size_t)(dh + 1) * dh * E); k->iw.b = rd((size_t)dh * E);
    k->iw.weight = rd((size_t)E * dh); k->iw.scaling = rd(E);
    if (model == 0) {
      k->ao_w = rd((size_t)d * d); k->ao_b = rd(d); k->ao_g = rd(d); k->ao_be = rd(d);
      k->ff1_w = rd((size_t)d * 2 * d); k->ff1_b = rd(2 * d); k->ff2_w = rd((size_t)2 * d * d); k->ff2_b = rd(d);
      k->ff_g = rd(d); k->ff_be = rd(d);
    } else {
      k->ln2_g = rd(d); k->ln2_b = rd(d);
      k->ff1_w = rd((size_t)d * d); k->ff1_b = rd(d); k->ff2_w = rd((size_t)d * d); k->ff2_b = rd(d);
    }
  }
  double *tr_w = NULL, *tr_b = NULL, *fin_g, *fin_b;
  if (model == 0) { tr_w = rd((size_t)d * d); tr_b = rd(d); }
  fin_g = rd(d); fin_b = rd(d);
  fclose(fin);

  FILE* fo = fopen(argv[2], "wb");
  const size_t LW = (size_t)L * W0;
  double* X = (double*)malloc(LW * sizeof(double));
  double* A = (double*)malloc(LW * sizeof(double));
  double* Q = (double*)malloc((size_t)L * 4 * d * sizeof(double));
  double *Kx = (double*)malloc((size_t)L * d * 8), *Vx = (double*)malloc((size_t)L * d * 8), *Tx = (double*)malloc((size_t)L * d * 8);
  double *O = (double*)malloc((size_t)L * d * 8), *P1 = (double*)malloc((size_t)L * 2 * d * 8), *P2 = (double*)malloc((size_t)L * 2 * d * 8);
  double *kmask = (double*)malloc(L * 8), *spans = (double*)malloc(L * 8), *logits = (double*)malloc((size_t)N1 * 8);
  int64_t* marks = (int64_t*)malloc((size_t)L * E * 8);
  for (int b = 0; b < B; ++b) {
    const int64_t* id = ids + (size_t)b * L;
    /* time: fp32 divides / subtract exactly like the fp32 graph (EasyDGL.py:71-74, CTSMA.py:47-49) */
    float tsf[4096];
    for (int l = 0; l < ts_len; ++l) tsf[l] = (float)ts[(size_t)b * ts_len + l] / (float)time_scale;
    for (int l = 0; l < L; ++l) {
      kmask[l] = id[l] != 0 ? 1.0 : 0.0;
      const int64_t mid = (model == 0 && id[l] == mask_id) ? 0 : id[l]; /* EasyDGL.py:76 */
      for (int e = 0; e < E; ++e) marks[(size_t)l * E + e] = mtab[(size_t)mid * E + e];
      if (model == 0) {
        const int l1 = l == 0 ? 1 : l;
        float sp = tsf[l1] - tsf[l1 - 1];
        sp = sp < 0.f ? 0.f : (sp > 100.f ? 100.f : sp); /* clip_by_value */
        spans[l] = sp;
      } else {
        spans[l] = (float)(tsf[l + 1] - tsf[l]);
      }
      for (int j = 0; j < d; ++j) X[(size_t)l * W0 + j] = item[(size_t)id[l] * d + j] * sqrt((double)d);
      if (model == 0) {
        for (int j = 0; j < d / 2; ++j) { /* coding.py:134-148 */
          const float sc = (float)pow(10000.0, (double)(2 * j) * 1.0 / (double)d);
          const float x = tsf[l] / sc;
          X[(size_t)l * W0 + 2 * j] += sin((double)x);
          X[(size_t)l * W0 + 2 * j + 1] += cos((double)x);
        }
        for (int j = 0; j < d; ++j) {
          X[(size_t)l * W0 + d + j] = pos[(size_t)l * d + j];
          double s = 0; /* EasyDGL.py:87-88: values index mark_embs */
          for (int e = 0; e < E; ++e) s += memb[(size_t)marks[(size_t)l * E + e] * d + j];
          X[(size_t)l * W0 + 2 * d + j] = s;
        }
      } else {
        for (int j = 0; j < d; ++j) X[(size_t)l * W0 + d + j] = pos[(size_t)l * d + j];
      }
    }
    int cin = W0;
    double* cur = X;
    for (int i = 0; i < nb; ++i) {
      Blk* k = &blk[i];
      if (model == 0) {
        dense(cur, cin, L, cin, k->qkvt_w, k->qkvt_b, 4 * d, Q); /* temporal.py:409 */
        for (int l = 0; l < L; ++l)
          for (int j = 0; j < d; ++j) {
            A[(size_t)l * d + j] = Q[(size_t)l * 4 * d + j];
            Kx[(size_t)l * d + j] = Q[(size_t)l * 4 * d + d + j];
            Vx[(size_t)l * d + j] = Q[(size_t)l * 4 * d + 2 * d + j];
            Tx[(size_t)l * d + j] = Q[(size_t)l * 4 * d + 3 * d + j];
          }
        attention(A, Kx, Vx, Tx, kmask, spans, marks, k->iw, L, d, h, E, 0, 1, O);
        for (int l = 0; l < L; ++l) for (int j = 0; j < d; ++j) O[(size_t)l * d + j] += cur[(size_t)l * cin + j]; /* :447 */
        dense(O, d, L, d, k->ao_w, k->ao_b, d, P1);                                                   /* EasyDGL.py:113 */
        for (int l = 0; l < L; ++l) for (int j = 0; j < d; ++j) P1[(size_t)l * d + j] += cur[(size_t)l * cin + j];
        layernorm(P1, L, d, k->ao_g, k->ao_be, A);                                                   /* :116 */
        dense(A, d, L, d, k->ff1_w, k->ff1_b, 2 * d, P2);                                            /* :120 */
        for (size_t t = 0; t < (size_t)L * 2 * d; ++t) P2[t] = gelu(P2[t]);
        dense(P2, 2 * d, L, 2 * d, k->ff2_w, k->ff2_b, d, P1);                                       /* :125 */
        for (size_t t = 0; t < (size_t)L * d; ++t) P1[t] += A[t];
        layernorm(P1, L, d, k->ff_g, k->ff_be, X);                                                   /* :128 */
      } else {
        layernorm(cur, L, cin, k->ln1_g, k->ln1_b, A);                                               /* CTSMA.py:68 */
        dense(A, cin, L, cin, k->q_w, k->q_b, d, Q);                                                 /* temporal.py:340-343 */
        dense(cur, cin, L, cin, k->k_w, k->k_b, d, Kx);
        dense(cur, cin, L, cin, k->v_w, k->v_b, d, Vx);
        dense(cur, cin, L, cin, k->t_w, k->t_b, d, Tx);
        attention(Q, Kx, Vx, Tx, kmask, spans, marks, k->iw, L, d, h, E, 1, 0, O);
        for (int l = 0; l < L; ++l) for (int j = 0; j < d; ++j) O[(size_t)l * d + j] += A[(size_t)l * cin + j]; /* :385 */
        layernorm(O, L, d, k->ln2_g, k->ln2_b, P1);                                                  /* CTSMA.py:73 */
        dense(P1, d, L, d, k->ff1_w, k->ff1_b, d, P2);                                               /* Base.py:79 */
        for (size_t t = 0; t < (size_t)L * d; ++t) P2[t] = P2[t] > 0 ? P2[t] : 0;
        dense(P2, d, L, d, k->ff2_w, k->ff2_b, d, X);                                                /* Base.py:83 */
        for (size_t t = 0; t < (size_t)L * d; ++t) X[t] += P1[t];                                    /* Base.py:86 */
      }
      cur = X;
      cin = d;
    }
    if (model == 0) {
      dense(cur, d, L, d, tr_w, tr_b, d, P1);                                                        /* EasyDGL.py:138 */
      for (size_t t = 0; t < (size_t)L * d; ++t) P1[t] = gelu(P1[t]);
      layernorm(P1, L, d, fin_g, fin_b, A);                                                          /* :139 */
    } else {
      layernorm(cur, L, d, fin_g, fin_b, A);                                                         /* CTSMA.py:80 */
    }
    const double* y = A + (size_t)(L - 1) * d;                                                       /* [:, -1] */
    for (int n = 0; n < N1; ++n) {
      double s = 0;
      for (int j = 0; j < d; ++j) s += y[j] * item[(size_t)n * d + j];
      logits[n] = s + (n == 0 ? -1000.0 : obias[n - 1]);                                             /* Base.py:110 */
    }
    fwrite(logits, 8, N1, fo);
  }
  fclose(fo);
  return 0;
}
