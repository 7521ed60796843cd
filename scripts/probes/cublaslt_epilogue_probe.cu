// Probe: does cuBLASLt offer the fused dense-layer epilogues for FP32 storage / TF32 compute on
// this GPU?  forward  y = relu(x w^T + b) + ReLU bit mask (RELU_AUX_BIAS)
//          backward dx = (dy w2) * mask, db = column sums (DRELU_BGRAD)
// Row-major tensors are handed to cuBLASLt as the column-major transposes.
// nvcc -arch=sm_100a -o probe cublaslt_epilogue_probe.cu -lcublasLt
#include <cublasLt.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { auto e_ = (x); if (e_ != 0) { printf("FAIL %s -> %d (line %d)\n", #x, (int)e_, __LINE__); return 1; } } while (0)

int main() {
  const int M = 4096, K = 256, N = 256, N2 = 44;
  std::vector<float> hx((size_t)M * K), hw((size_t)N * K), hb(N), hdy((size_t)M * N2), hw2((size_t)N2 * N);
  srand(1);
  auto rnd = [] { return (rand() % 2001 - 1000) / 1000.0f; };
  for (auto &v : hx) v = rnd();
  for (auto &v : hw) v = rnd() * 0.1f;
  for (auto &v : hb) v = rnd();
  for (auto &v : hdy) v = rnd();
  for (auto &v : hw2) v = rnd() * 0.1f;
  float *x, *w, *b, *y, *dy, *w2, *dx, *db;
  unsigned char *mask;
  CK(cudaMalloc(&x, hx.size() * 4)); CK(cudaMalloc(&w, hw.size() * 4)); CK(cudaMalloc(&b, N * 4));
  CK(cudaMalloc(&y, (size_t)M * N * 4)); CK(cudaMalloc(&dy, hdy.size() * 4));
  CK(cudaMalloc(&w2, hw2.size() * 4)); CK(cudaMalloc(&dx, (size_t)M * N * 4)); CK(cudaMalloc(&db, N * 4));
  CK(cudaMalloc(&mask, (size_t)M * N / 8));
  CK(cudaMemcpy(x, hx.data(), hx.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(w, hw.data(), hw.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(b, hb.data(), N * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dy, hdy.data(), hdy.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(w2, hw2.data(), hw2.size() * 4, cudaMemcpyHostToDevice));
  void *ws; const size_t ws_bytes = 64 << 20; CK(cudaMalloc(&ws, ws_bytes));
  cublasLtHandle_t lt; CK(cublasLtCreate(&lt));
  for (int tf32 = 0; tf32 < 2; tf32++) {
    const cublasComputeType_t ct = tf32 ? CUBLAS_COMPUTE_32F_FAST_TF32 : CUBLAS_COMPUTE_32F;
    // ---- forward: D[N, M] = w^T-view^T[N, K] * x-view[K, M]
    {
      cublasLtMatmulDesc_t op; CK(cublasLtMatmulDescCreate(&op, ct, CUDA_R_32F));
      cublasOperation_t ta = CUBLAS_OP_T, tb = CUBLAS_OP_N;
      CK(cublasLtMatmulDescSetAttribute(op, CUBLASLT_MATMUL_DESC_TRANSA, &ta, sizeof(ta)));
      CK(cublasLtMatmulDescSetAttribute(op, CUBLASLT_MATMUL_DESC_TRANSB, &tb, sizeof(tb)));
      cublasLtEpilogue_t ep = CUBLASLT_EPILOGUE_RELU_AUX_BIAS;
      CK(cublasLtMatmulDescSetAttribute(op, CUBLASLT_MATMUL_DESC_EPILOGUE, &ep, sizeof(ep)));
      CK(cublasLtMatmulDescSetAttribute(op, CUBLASLT_MATMUL_DESC_BIAS_POINTER, &b, sizeof(b)));
      CK(cublasLtMatmulDescSetAttribute(op, CUBLASLT_MATMUL_DESC_EPILOGUE_AUX_POINTER, &mask, sizeof(mask)));
      int64_t ld = N; CK(cublasLtMatmulDescSetAttribute(op, CUBLASLT_MATMUL_DESC_EPILOGUE_AUX_LD, &ld, sizeof(ld)));
      cublasLtMatrixLayout_t la, lb, ld_;
      CK(cublasLtMatrixLayoutCreate(&la, CUDA_R_32F, K, N, K));   // w buffer: col-major K x N
      CK(cublasLtMatrixLayoutCreate(&lb, CUDA_R_32F, K, M, K));   // x buffer: col-major K x M
      CK(cublasLtMatrixLayoutCreate(&ld_, CUDA_R_32F, N, M, N));  // y buffer: col-major N x M
      cublasLtMatmulPreference_t pref; CK(cublasLtMatmulPreferenceCreate(&pref));
      CK(cublasLtMatmulPreferenceSetAttribute(pref, CUBLASLT_MATMUL_PREF_MAX_WORKSPACE_BYTES, &ws_bytes, sizeof(ws_bytes)));
      cublasLtMatmulHeuristicResult_t res[4]; int n = 0;
      auto st = cublasLtMatmulAlgoGetHeuristic(lt, op, la, lb, ld_, ld_, pref, 4, res, &n);
      printf("forward RELU_AUX_BIAS tf32=%d: heuristic status %d, %d algos\n", tf32, (int)st, n);
      if (st == 0 && n > 0) {
        float one = 1.f, zero = 0.f;
        CK(cublasLtMatmul(lt, op, &one, w, la, x, lb, &zero, y, ld_, y, ld_, &res[0].algo, ws, ws_bytes, 0));
        CK(cudaDeviceSynchronize());
        std::vector<float> hy((size_t)M * N); CK(cudaMemcpy(hy.data(), y, hy.size() * 4, cudaMemcpyDeviceToHost));
        double maxerr = 0;
        for (int r = 0; r < 64; r++) for (int j = 0; j < N; j++) {
          double acc = hb[j];
          for (int k = 0; k < K; k++) acc += (double)hx[(size_t)r * K + k] * hw[(size_t)j * K + k];
          const double ref = acc > 0 ? acc : 0;
          maxerr = fmax(maxerr, fabs(ref - hy[(size_t)r * N + j]));
        }
        printf("  forward max |err| over 64 rows: %.3e\n", maxerr);
      }
    }
    // ---- backward: D[N, M] = w2-view[N, N2] * dy-view[N2, M], dReLU with the mask, bgrad
    {
      cublasLtMatmulDesc_t op; CK(cublasLtMatmulDescCreate(&op, ct, CUDA_R_32F));
      cublasOperation_t ta = CUBLAS_OP_N, tb = CUBLAS_OP_N;
      CK(cublasLtMatmulDescSetAttribute(op, CUBLASLT_MATMUL_DESC_TRANSA, &ta, sizeof(ta)));
      CK(cublasLtMatmulDescSetAttribute(op, CUBLASLT_MATMUL_DESC_TRANSB, &tb, sizeof(tb)));
      cublasLtEpilogue_t ep = CUBLASLT_EPILOGUE_DRELU_BGRAD;
      CK(cublasLtMatmulDescSetAttribute(op, CUBLASLT_MATMUL_DESC_EPILOGUE, &ep, sizeof(ep)));
      CK(cublasLtMatmulDescSetAttribute(op, CUBLASLT_MATMUL_DESC_BIAS_POINTER, &db, sizeof(db)));
      CK(cublasLtMatmulDescSetAttribute(op, CUBLASLT_MATMUL_DESC_EPILOGUE_AUX_POINTER, &mask, sizeof(mask)));
      int64_t ld = N; CK(cublasLtMatmulDescSetAttribute(op, CUBLASLT_MATMUL_DESC_EPILOGUE_AUX_LD, &ld, sizeof(ld)));
      cublasLtMatrixLayout_t la, lb, ld_;
      CK(cublasLtMatrixLayoutCreate(&la, CUDA_R_32F, N, N2, N));   // w2 buffer [N2, N] row-major = col-major N x N2
      CK(cublasLtMatrixLayoutCreate(&lb, CUDA_R_32F, N2, M, N2));  // dy buffer: col-major N2 x M
      CK(cublasLtMatrixLayoutCreate(&ld_, CUDA_R_32F, N, M, N));   // dx buffer: col-major N x M
      cublasLtMatmulPreference_t pref; CK(cublasLtMatmulPreferenceCreate(&pref));
      CK(cublasLtMatmulPreferenceSetAttribute(pref, CUBLASLT_MATMUL_PREF_MAX_WORKSPACE_BYTES, &ws_bytes, sizeof(ws_bytes)));
      cublasLtMatmulHeuristicResult_t res[4]; int n = 0;
      auto st = cublasLtMatmulAlgoGetHeuristic(lt, op, la, lb, ld_, ld_, pref, 4, res, &n);
      printf("backward DRELU_BGRAD tf32=%d: heuristic status %d, %d algos\n", tf32, (int)st, n);
      if (st == 0 && n > 0) {
        float one = 1.f, zero = 0.f;
        CK(cublasLtMatmul(lt, op, &one, w2, la, dy, lb, &zero, dx, ld_, dx, ld_, &res[0].algo, ws, ws_bytes, 0));
        CK(cudaDeviceSynchronize());
        std::vector<float> hdx((size_t)M * N), hy((size_t)M * N), hdb(N);
        CK(cudaMemcpy(hdx.data(), dx, hdx.size() * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(hy.data(), y, hy.size() * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(hdb.data(), db, N * 4, cudaMemcpyDeviceToHost));
        double maxerr = 0, maxdb = 0;
        std::vector<double> col(N, 0.0);
        for (int r = 0; r < M; r++) for (int j = 0; j < N; j++) {
          double acc = 0;
          if (r < 64 || true) for (int k = 0; k < N2; k++) acc += (double)hdy[(size_t)r * N2 + k] * hw2[(size_t)k * N + j];
          const double ref = hy[(size_t)r * N + j] > 0 ? acc : 0;
          col[j] += ref;
          if (r < 64) maxerr = fmax(maxerr, fabs(ref - hdx[(size_t)r * N + j]));
        }
        for (int j = 0; j < N; j++) maxdb = fmax(maxdb, fabs(col[j] - hdb[j]));
        printf("  backward max |err| dx (64 rows): %.3e, db: %.3e\n", maxerr, maxdb);
      }
    }
  }
  return 0;
}
