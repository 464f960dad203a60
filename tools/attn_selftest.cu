// attn_selftest.cu - torch-free GPU self-test of the attention kernels (attn_tc2.cu vs the CUDA-core kernel and a
// double-precision host restatement of temporal.py:281-315, 345-385, 412-447), with per-phase intermediates of the
// tcgen05 kernel and a timing of the C2 shape.  Test infrastructure only.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tools/attn_selftest tools/attn_selftest.cu \
//        -Leasydgl_b200/csrc -leasydgl_b200 -Xlinker -rpath -Xlinker '$ORIGIN/../easydgl_b200/csrc'
//   tools/attn_selftest [B h L causal diag [bench]]
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <random>
#include <vector>

#include "../easydgl_b200/csrc/common.cuh"

using namespace edgl;

#define CK(x)                                                                                   \
  do {                                                                                          \
    cudaError_t e_ = (x);                                                                       \
    if (e_ != cudaSuccess) {                                                                    \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);           \
      exit(2);                                                                                  \
    }                                                                                           \
  } while (0)

template <typename T>
T* dev(const std::vector<T>& v) {
  T* p = nullptr;
  CK(cudaMalloc(&p, v.size() * sizeof(T) + 16));
  CK(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return p;
}

int main(int argc, char** argv) {
  const int B = argc > 1 ? atoi(argv[1]) : 3, h = argc > 2 ? atoi(argv[2]) : 2, L = argc > 3 ? atoi(argv[3]) : 100;
  const bool causal = argc > 4 ? atoi(argv[4]) != 0 : false, diag = argc > 5 ? atoi(argv[5]) != 0 : true;
  const bool bench = argc > 6 ? atoi(argv[6]) != 0 : false;
  const int DH = 16, E = 16, d = DH * h, NC = DH * E;
  const long long rows = (long long)B * L;
  std::mt19937 rng(1234);
  std::normal_distribution<float> nd(0.f, 1.f);
  std::uniform_real_distribution<float> ud(0.f, 1.f);
  std::vector<float> qkvt((size_t)rows * 4 * d), spans(rows), R((size_t)rows * d), w1(17 * NC), b1(NC), wv(NC), scal(E);
  std::vector<uint8_t> kmask(rows), marks((size_t)rows * E);
  for (auto& x : qkvt) x = nd(rng);
  for (auto& x : R) x = nd(rng);
  for (auto& x : spans) x = 100.f * ud(rng);
  for (auto& x : w1) x = ud(rng) - 0.5f;
  for (auto& x : b1) x = 0.3f * nd(rng);
  for (auto& x : wv) x = 1.5f * (ud(rng) - 0.5f);
  for (auto& x : scal) x = ud(rng) - 0.5f;
  for (int i = 0; i < NC; ++i) w1[16 * NC + i] *= 0.02f;  // span row
  for (int b = 0; b < B; ++b) {
    const int pad = (b == 1) ? L : (int)(ud(rng) * (L / 2));  // sequence 1: all keys masked (uniform attention, Q8)
    for (int l = 0; l < L; ++l) {
      kmask[(size_t)b * L + l] = l >= pad ? 1 : 0;
      for (int e = 0; e < E; ++e) marks[((size_t)b * L + l) * E + e] = 0;
      const int nm = 1 + (int)(ud(rng) * 3);
      for (int j = 0; j < nm; ++j) marks[((size_t)b * L + l) * E + (int)(ud(rng) * E) % E] = 1;
    }
  }
  float *dq = dev(qkvt), *dsp = dev(spans), *dR = dev(R), *dw1 = dev(w1), *db1 = dev(b1), *dwv = dev(wv), *dsc = dev(scal);
  uint8_t *dkm = dev(kmask), *dmk = dev(marks);
  float *dO1, *dO2, *dl1, *dl2, *ddbg = nullptr;
  CK(cudaMalloc(&dO1, rows * d * 4)); CK(cudaMalloc(&dO2, rows * d * 4));
  CK(cudaMalloc(&dl1, (size_t)h * rows * E * 4)); CK(cudaMalloc(&dl2, (size_t)h * rows * E * 4));
  CK(cudaMemset(dO1, 0, rows * d * 4)); CK(cudaMemset(dO2, 0, rows * d * 4));
  void* pack = nullptr;
  CK(cudaMalloc(&pack, attention_tc2_pack_bytes(DH, E)));
  if (launch_attention_tc2_pack(dw1, db1, dwv, dsc, DH, E, pack, 0)) { printf("pack: %s\n", last_error().c_str()); return 2; }
  const int items = B * h;
  const size_t dbg_n = (size_t)items * 8 * 128 * 256;
  const bool want_dbg = !bench && items <= 16;
  if (want_dbg) { CK(cudaMalloc(&ddbg, dbg_n * 4)); CK(cudaMemset(ddbg, 0, dbg_n * 4)); }

  AttnArgs a;
  memset(&a, 0, sizeof(a));
  a.Q = dq; a.K = dq + d; a.V = dq + 2 * d; a.T = dq + 3 * d;
  a.ldq = a.ldk = a.ldv = a.ldt = 4 * d;
  a.kmask = dkm; a.spans = dsp; a.marks = dmk; a.R = dR; a.ldr = d;
  a.int_w = dw1; a.int_b = db1; a.int_weight = dwv; a.int_scaling = dsc;
  a.O = dO1; a.ldo = d; a.lam = dl1; a.B = B; a.L = L; a.d = d; a.h = h; a.E = E;
  a.causal = causal; a.diag_one = diag;
  a.mlp_pack2 = pack;

  if (bench) {
    // timing: tc2 vs the mma.sync f16 kernel is done from python (the f16 kernel needs its own pack); here tc2 only
    a.lam = nullptr;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int it = 0; it < 3; ++it)
      if (launch_attention_mode(a, 0, '2')) { printf("tc2: %s\n", last_error().c_str()); return 2; }
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    const int iters = 20;
    for (int it = 0; it < iters; ++it) launch_attention_mode(a, 0, '2');
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("BENCH tc2 B=%d h=%d L=%d causal=%d: %.4f ms per launch\n", B, h, L, (int)causal, ms / iters);
    // per-phase cycle counters of the first thread of every slot (PROF instantiation)
    long long* dprof = nullptr;
    const int nslots = 148 * 2;
    CK(cudaMalloc(&dprof, nslots * 16 * sizeof(long long)));
    CK(cudaMemset(dprof, 0, nslots * 16 * sizeof(long long)));
    a.prof = dprof;
    launch_attention_mode(a, 0, '2');
    CK(cudaDeviceSynchronize());
    std::vector<long long> pr(nslots * 16);
    CK(cudaMemcpy(pr.data(), dprof, pr.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    const char* pn[14] = {"wait S", "softmax pass1+bar", "softmax pass2+st+bar", "wait PT", "H conv+bar", "wait MLP", "sigmoids",
                          "lam st+bar", "wait G", "gate+bar", "convert next", "wait PV+bar", "epilogue", "loop head"};
    double tot = 0;
    double per[14];
    const double items_per_slot = (double)B * h / nslots;
    for (int i = 0; i < 14; ++i) {
      double sum = 0;
      for (int sidx = 0; sidx < nslots; ++sidx) sum += (double)pr[sidx * 16 + i];
      per[i] = sum / nslots / items_per_slot;
      tot += per[i];
    }
    for (int i = 0; i < 14; ++i) printf("  PROF %-22s %8.0f cycles/item (%4.1f %%)\n", pn[i], per[i], 100.0 * per[i] / tot);
    printf("  PROF total %.0f cycles per item and slot\n", tot);
    return 0;
  }

  // ---- reference kernel (exact-fp32 CUDA cores) and the kernel under test
  if (launch_attention_mode(a, 0, 's')) { printf("simt: %s\n", last_error().c_str()); return 2; }
  CK(cudaDeviceSynchronize());
  AttnArgs t = a;
  t.O = dO2; t.lam = dl2; t.dbg = ddbg;
  const int rc = launch_attention_tc2(t, 0);
  if (rc != 0) { printf("tc2 rc=%d: %s\n", rc, last_error().c_str()); return 2; }
  cudaError_t se = cudaDeviceSynchronize();
  if (se != cudaSuccess) { printf("tc2 kernel failed: %s\n", cudaGetErrorString(se)); return 2; }
  std::vector<float> O1((size_t)rows * d), O2((size_t)rows * d), l1((size_t)h * rows * E), l2((size_t)h * rows * E);
  CK(cudaMemcpy(O1.data(), dO1, O1.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(O2.data(), dO2, O2.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(l1.data(), dl1, l1.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(l2.data(), dl2, l2.size() * 4, cudaMemcpyDeviceToHost));
  std::vector<float> dbg;
  if (want_dbg) { dbg.resize(dbg_n); CK(cudaMemcpy(dbg.data(), ddbg, dbg_n * 4, cudaMemcpyDeviceToHost)); }

  // ---- double-precision host reference with intermediates
  const double fill = -4294967296.0, log2e = 1.4426950408889634;
  double eO = 0, mO = 0, eOs = 0, eL = 0, mL = 0;
  double ph_err[7] = {0, 0, 0, 0, 0, 0, 0}, ph_max[7] = {0, 0, 0, 0, 0, 0, 0};
  std::vector<double> S(L), P(L), G(L);
  for (int b = 0; b < B; ++b)
    for (int hh = 0; hh < h; ++hh) {
      const int item = b * h + hh;
      for (int q = 0; q < L; ++q) {
        const float* Q = &qkvt[((size_t)b * L + q) * 4 * d + hh * DH];
        double m = -1e300;
        for (int k = 0; k < L; ++k) {
          const float* K = &qkvt[((size_t)b * L + k) * 4 * d + d + hh * DH];
          double s = 0;
          for (int j = 0; j < DH; ++j) s += (double)Q[j] * K[j];
          s /= 4.0;
          if (!kmask[(size_t)b * L + k]) s = fill;
          if (causal && k > q) s = fill;
          S[k] = s;
          m = fmax(m, s);
        }
        double l = 0;
        for (int k = 0; k < L; ++k) { P[k] = exp(S[k] - m); l += P[k]; }
        for (int k = 0; k < L; ++k) P[k] /= l;
        double H[16] = {0};
        for (int k = 0; k < L; ++k) {
          const float* T = &qkvt[((size_t)b * L + k) * 4 * d + 3 * d + hh * DH];
          for (int j = 0; j < DH; ++j) H[j] += P[k] * T[j];
        }
        double lam[16], x[16];
        const double sp = spans[(size_t)b * L + q];
        for (int e = 0; e < E; ++e) {
          double acc = 0;
          for (int j = 0; j < DH; ++j) {
            const int c = e * DH + j;
            double z = b1[c] + sp * w1[16 * NC + c];
            for (int i = 0; i < DH; ++i) z += H[i] * w1[i * NC + c];
            acc += wv[c] / (1.0 + exp(-z));
          }
          x[e] = acc;
          const double s = exp((double)scal[e]);
          lam[e] = s * log(1.0 + exp(acc / s));
        }
        double O[16] = {0};
        for (int k = 0; k < L; ++k) {
          double g = 0;
          for (int e = 0; e < E; ++e) g += lam[e] * marks[((size_t)b * L + k) * E + e];
          if (diag && k == q) g = 1.0;
          G[k] = g;
          const float* V = &qkvt[((size_t)b * L + k) * 4 * d + 2 * d + hh * DH];
          for (int j = 0; j < DH; ++j) O[j] += g * P[k] * V[j];
        }
        for (int j = 0; j < DH; ++j) {
          const double ref = O[j] + R[((size_t)b * L + q) * d + hh * DH + j];
          const size_t o = ((size_t)b * L + q) * d + hh * DH + j;
          eO = fmax(eO, fabs(O2[o] - ref)); eOs = fmax(eOs, fabs(O1[o] - ref)); mO = fmax(mO, fabs(ref));
        }
        for (int e = 0; e < E; ++e) {
          const size_t o = (((size_t)hh * B + b) * L + q) * E + e;
          eL = fmax(eL, fabs(l2[o] - lam[e])); mL = fmax(mL, fabs(lam[e]));
        }
        if (want_dbg) {
          auto D = [&](int phase, int col) { return (double)dbg[(((size_t)item * 8 + phase) * 128 + q) * 256 + col]; };
          auto upd = [&](int phase, double got, double ref) {
            if (!(fabs(ref) > 1e200)) { ph_err[phase] = fmax(ph_err[phase], fabs(got - ref)); ph_max[phase] = fmax(ph_max[phase], fabs(ref)); }
          };
          for (int k = 0; k < L; ++k) {
            if (S[k] > -1e9) upd(0, D(0, k), S[k] * log2e);
            upd(1, D(1, k), P[k]);
            upd(5, D(5, k), G[k]);
          }
          for (int j = 0; j < DH; ++j) { upd(2, D(2, j), H[j]); upd(6, D(6, j), O[j] + R[((size_t)b * L + q) * d + hh * DH + j]); }
          for (int e = 0; e < E; ++e) { upd(3, D(3, e), x[e]); upd(4, D(4, e), lam[e]); }
        }
      }
    }
  printf("shape B=%d h=%d L=%d causal=%d diag=%d\n", B, h, L, (int)causal, (int)diag);
  printf("simt vs double: out %.3e (max|ref| %.3e)\n", eOs / mO, mO);
  printf("tc2  vs double: out %.3e  lam %.3e\n", eO / mO, eL / mL);
  if (want_dbg) {
    const char* nm[7] = {"S(log2)", "P", "H", "x", "lam", "G", "O"};
    for (int p = 0; p < 7; ++p) printf("  phase %-8s rel err %.3e (max|ref| %.3e)\n", nm[p], ph_err[p] / fmax(ph_max[p], 1e-300), ph_max[p]);
  }
  const bool ok = eO / mO < 2e-5 && eL / mL < 2e-5;
  printf(ok ? "SELFTEST_OK\n" : "SELFTEST_FAILED\n");
  return ok ? 0 : 1;
}
