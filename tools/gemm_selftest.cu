// gemm_selftest.cu - a torch-free check of the tcgen05 dense entry points through the C ABI (dlopen), for boxes where
// only seconds of GPU time are available: edgl_dense_nk (3xTF32) and edgl_dense_nk_f16 (scaled 3xFP16) against a
// double-precision CPU product, bias + GELU / ReLU / none, both tile widths, M / N / K tails.
//   nvcc -O2 -o tools/gemm_selftest tools/gemm_selftest.cu -ldl && EDGL_TC_EPI=staged tools/gemm_selftest && tools/gemm_selftest
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

typedef int (*dense_fn)(const float*, const float*, const float*, int, int, int, int, float*, void*);
typedef const char* (*err_fn)(void);

static double gelu(double x) { return 0.5 * x * (1.0 + erf(x / sqrt(2.0))); }

int main(int argc, char** argv) {
  const char* path = argc > 1 ? argv[1] : "easydgl_b200/csrc/libeasydgl_b200.so";
  void* lib = dlopen(path, RTLD_NOW);
  if (!lib) { printf("dlopen failed: %s\n", dlerror()); return 2; }
  dense_fn fns[2] = {(dense_fn)dlsym(lib, "edgl_dense_nk"), (dense_fn)dlsym(lib, "edgl_dense_nk_f16")};
  err_fn last_error = (err_fn)dlsym(lib, "edgl_last_error");
  const char* names[2] = {"3xTF32", "3xFP16"};
  const int shapes[][4] = {{300, 128, 256, 1}, {257, 144, 512, 0}, {1000, 256, 128, 1}, {130, 64, 128, 2}, {77, 128, 1001, 0}};
  int bad = 0;
  srand(7);
  for (auto& sh : shapes) {
    const int M = sh[0], K = sh[1], N = sh[2], act = sh[3];
    std::vector<float> x((size_t)M * K), wt((size_t)N * K), b(N), out((size_t)M * N);
    for (auto& v : x) v = (float)rand() / RAND_MAX * 2.f - 1.f;
    for (auto& v : wt) v = ((float)rand() / RAND_MAX * 2.f - 1.f) / sqrtf((float)K);
    for (auto& v : b) v = ((float)rand() / RAND_MAX * 2.f - 1.f) * 0.1f;
    std::vector<double> ref((size_t)M * N);
    double rmax = 0;
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < N; ++n) {
        double s = b[n];
        for (int k = 0; k < K; ++k) s += (double)x[(size_t)m * K + k] * wt[(size_t)n * K + k];
        if (act == 1) s = gelu(s);
        if (act == 2) s = s > 0 ? s : 0;
        ref[(size_t)m * N + n] = s;
        rmax = fmax(rmax, fabs(s));
      }
    float *dx, *dw, *db, *dout;
    cudaMalloc(&dx, x.size() * 4); cudaMalloc(&dw, wt.size() * 4); cudaMalloc(&db, b.size() * 4); cudaMalloc(&dout, out.size() * 4);
    cudaMemcpy(dx, x.data(), x.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dw, wt.data(), wt.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(db, b.data(), b.size() * 4, cudaMemcpyHostToDevice);
    for (int f = 0; f < 2; ++f) {
      cudaMemset(dout, 0xff, out.size() * 4);
      const int rc = fns[f](dx, dw, db, M, K, N, act, dout, nullptr);
      cudaError_t ce = cudaDeviceSynchronize();
      cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
      double emax = 0;
      for (size_t i = 0; i < out.size(); ++i) {
        const double e = fabs((double)out[i] - ref[i]);
        if (!(e <= emax)) emax = e;  // also catches NaN
      }
      const bool ok = rc == 0 && ce == cudaSuccess && emax <= 1e-5 * rmax;
      printf("%s M=%d K=%d N=%d act=%d rc=%d cuda=%d max|err|/max|ref| = %.3e %s\n", names[f], M, K, N, act, rc, (int)ce,
             emax / rmax, ok ? "ok" : "FAIL");
      if (rc) printf("  %s\n", last_error());
      bad += !ok;
    }
    cudaFree(dx); cudaFree(dw); cudaFree(db); cudaFree(dout);
  }
  printf(bad ? "SELFTEST_FAILED\n" : "SELFTEST_OK\n");
  return bad ? 1 : 0;
}
