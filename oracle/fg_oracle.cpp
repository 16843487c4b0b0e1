// fg_oracle.cpp -- CPU ORACLE for the aleju/face-generator GAN train-step hot path.
//
// THIS IS TEST INFRASTRUCTURE, NOT PRODUCT.  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load it.  The product path
// (face_generator_b200/csrc/*.cu behind include/fg_b200.h) never calls into this file.
//
// PARITY UNPINNED: the reference (Lua/Torch7) ships no tests, golden vectors or fixtures
// and cannot be executed in this image (no LuaJIT/Torch7); the arithmetic lives in the
// un-vendored, un-pinned `nn`/`cunn`/`cudnn.torch` rocks of Dec-2015/Jan-2016.  This file
// restates that published algorithm family (THNN: per-sample im2col + GEMM convolution,
// batch statistics BN, shared-slope PReLU, ...) and is cross-checked in
// tests/test_oracle_vs_torch.py against PyTorch-CPU (same THNN lineage).
//
// Layouts are the reference's: fp32-style dense NCHW activations, conv weights
// [Cout][Cin][kH][kW], Linear weights [out][in], flat parameter vectors in
// `getParameters()` order (module order, weight then bias).
//
// Two instantiations of the same template code are exported:
//   *_f64 : double storage + double accumulation  -> the parity oracle
//   *_f32 : float storage + float accumulation, OpenMP -> the CPU baseline "port"
//           (THNN's algorithm: per-sample im2col + SGEMM, batch-parallel)
//
// Reference citations (file:line under /root/reference):
//   G topology            models.lua:57-81   (create_G_decoder_upsampling32)
//   D topology            models.lua:382-416 (create_D32b)
//   train loop body       adversarial.lua:54-300 (fevalD :83-179, fevalG_on_D :187-231,
//                         batch assembly :240-257, G step :275-288)
//   Adam                  interruptable_optimizers.lua:49-94
//   noise                 utils/nn_utils.lua:35-39
//   SCU (c2f, factor=1)   layers/SpatialConvolutionUpsample.lua:13-30
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <type_traits>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// Optional SGEMM for the fp32 "port" (the CPU baseline): Torch7's nn calls the system BLAS (OpenBLAS / MKL) for the
// per-sample GEMM of SpatialConvolutionMM and for nn.Linear, so a fair reference-style baseline does too.
// fgo_use_blas() dlopens an OpenBLAS (the one bundled with scipy in this image); without it the blocked loops below
// are used.  The fp64 parity oracle never takes this path.
typedef void (*sgemm_fn)(int order, int ta, int tb, int M, int N, int K, float alpha, const float* A, int lda,
                         const float* B, int ldb, float beta, float* C, int ldc);
sgemm_fn g_sgemm = nullptr;
constexpr int kRowMajor = 101, kNoTrans = 111, kTrans = 112;

// ----------------------------------------------------------------------------------------------
// dense helpers
// ----------------------------------------------------------------------------------------------
// C[M,N] (+)= A[M,K] * B[K,N]            (row-major, "NN")
template <class T>
void gemm_nn(int M, int N, int K, const T* A, const T* B, T* C, bool acc) {
  int i = 0;
  // 4 output rows at a time share every streamed row of B (4x less memory traffic than one row at a time)
  for (; i + 3 < M; i += 4) {
    T* c0 = C + (size_t)i * N;
    T *c1 = c0 + N, *c2 = c1 + N, *c3 = c2 + N;
    if (!acc) std::fill(c0, c0 + (size_t)4 * N, T(0));
    const T* a0 = A + (size_t)i * K;
    const T *a1 = a0 + K, *a2 = a1 + K, *a3 = a2 + K;
    for (int k = 0; k < K; ++k) {
      const T v0 = a0[k], v1 = a1[k], v2 = a2[k], v3 = a3[k];
      const T* b = B + (size_t)k * N;
#pragma omp simd
      for (int j = 0; j < N; ++j) {
        const T bj = b[j];
        c0[j] += v0 * bj;
        c1[j] += v1 * bj;
        c2[j] += v2 * bj;
        c3[j] += v3 * bj;
      }
    }
  }
  for (; i < M; ++i) {
    T* c = C + (size_t)i * N;
    if (!acc) std::fill(c, c + N, T(0));
    const T* a = A + (size_t)i * K;
    for (int k = 0; k < K; ++k) {
      const T av = a[k];
      const T* b = B + (size_t)k * N;
#pragma omp simd
      for (int j = 0; j < N; ++j) c[j] += av * b[j];
    }
  }
}
// C[M,N] (+)= A[K,M]^T * B[K,N]           ("TN")
template <class T>
void gemm_tn(int M, int N, int K, const T* A, const T* B, T* C, bool acc) {
  if (!acc) std::fill(C, C + (size_t)M * N, T(0));
  for (int k = 0; k < K; ++k) {
    const T* a = A + (size_t)k * M;
    const T* b = B + (size_t)k * N;
    for (int i = 0; i < M; ++i) {
      const T av = a[i];
      T* c = C + (size_t)i * N;
#pragma omp simd
      for (int j = 0; j < N; ++j) c[j] += av * b[j];
    }
  }
}
// C[M,N] (+)= A[M,K] * B[N,K]^T           ("NT")
template <class T>
void gemm_nt(int M, int N, int K, const T* A, const T* B, T* C, bool acc) {
  int i = 0;
  for (; i + 3 < M; i += 4) {  // 4 rows of A share each streamed row of B
    const T* a0 = A + (size_t)i * K;
    const T *a1 = a0 + K, *a2 = a1 + K, *a3 = a2 + K;
    for (int j = 0; j < N; ++j) {
      const T* b = B + (size_t)j * K;
      T s0 = 0, s1 = 0, s2 = 0, s3 = 0;
#pragma omp simd reduction(+ : s0, s1, s2, s3)
      for (int k = 0; k < K; ++k) {
        const T bk = b[k];
        s0 += a0[k] * bk;
        s1 += a1[k] * bk;
        s2 += a2[k] * bk;
        s3 += a3[k] * bk;
      }
      T* c = C + (size_t)i * N + j;
      c[0] = acc ? c[0] + s0 : s0;
      c[(size_t)N] = acc ? c[(size_t)N] + s1 : s1;
      c[(size_t)2 * N] = acc ? c[(size_t)2 * N] + s2 : s2;
      c[(size_t)3 * N] = acc ? c[(size_t)3 * N] + s3 : s3;
    }
  }
  for (; i < M; ++i) {
    const T* a = A + (size_t)i * K;
    for (int j = 0; j < N; ++j) {
      const T* b = B + (size_t)j * K;
      T s = 0;
#pragma omp simd reduction(+ : s)
      for (int k = 0; k < K; ++k) s += a[k] * b[k];
      T& c = C[(size_t)i * N + j];
      c = acc ? c + s : s;
    }
  }
}

// ----------------------------------------------------------------------------------------------
// nn.Linear: y = x W^T + b ; W[out][in]            (models.lua:59, :406-412)
// ----------------------------------------------------------------------------------------------
template <class T>
void linear_fwd(int B, int in, int out, const T* x, const T* W, const T* b, T* y) {
  if constexpr (std::is_same<T, float>::value) {
    if (g_sgemm) {  // y = x W^T + b   (nn.Linear:updateOutput: addmm + addr)
      g_sgemm(kRowMajor, kNoTrans, kTrans, B, out, in, 1.f, x, in, W, in, 0.f, y, out);
      for (int n = 0; n < B; ++n)
        for (int j = 0; j < out; ++j) y[(size_t)n * out + j] += b[j];
      return;
    }
  }
#pragma omp parallel for schedule(static)
  for (int n = 0; n < B; ++n) {
    gemm_nt(1, out, in, x + (size_t)n * in, W, y + (size_t)n * out, false);
    for (int j = 0; j < out; ++j) y[(size_t)n * out + j] += b[j];
  }
}
// dx = dy W ; dW += dy^T x ; db += sum_n dy
template <class T>
void linear_bwd(int B, int in, int out, const T* x, const T* W, const T* dy, T* dx, T* dW, T* db) {
  if constexpr (std::is_same<T, float>::value) {
    if (g_sgemm) {
      if (dx) g_sgemm(kRowMajor, kNoTrans, kNoTrans, B, in, out, 1.f, dy, out, W, in, 0.f, dx, in);   // dx = dy W
      if (dW) g_sgemm(kRowMajor, kTrans, kNoTrans, out, in, B, 1.f, dy, out, x, in, 1.f, dW, in);     // dW += dy^T x
      if (db)
        for (int j = 0; j < out; ++j) {
          float s = 0;
          for (int n = 0; n < B; ++n) s += dy[(size_t)n * out + j];
          db[j] += s;
        }
      return;
    }
  }
  if (dx) {
#pragma omp parallel for schedule(static)
    for (int n = 0; n < B; ++n) gemm_nn(1, in, out, dy + (size_t)n * out, W, dx + (size_t)n * in, false);
  }
  if (dW) {
#pragma omp parallel for schedule(static)
    for (int j = 0; j < out; ++j) {
      T* w = dW + (size_t)j * in;
      for (int n = 0; n < B; ++n) {
        const T g = dy[(size_t)n * out + j];
        const T* xr = x + (size_t)n * in;
#pragma omp simd
        for (int k = 0; k < in; ++k) w[k] += g * xr[k];
      }
    }
  }
  if (db) {
    for (int j = 0; j < out; ++j) {
      T s = 0;
      for (int n = 0; n < B; ++n) s += dy[(size_t)n * out + j];
      db[j] += s;
    }
  }
}

// ----------------------------------------------------------------------------------------------
// SpatialConvolution, stride 1, pad (k-1)/2, cross-correlation, NCHW  (models.lua:64,69,73,385-400)
// THNN SpatialConvolutionMM algorithm: per-sample im2col + GEMM.
// ----------------------------------------------------------------------------------------------
template <class T>
void im2col(const T* x, int Cin, int H, int W, int k, T* col) {
  const int pad = (k - 1) / 2;
  for (int c = 0; c < Cin; ++c)
    for (int kh = 0; kh < k; ++kh)
      for (int kw = 0; kw < k; ++kw) {
        T* dst = col + ((size_t)(c * k + kh) * k + kw) * H * W;
        for (int h = 0; h < H; ++h) {
          const int ih = h + kh - pad;
          if (ih < 0 || ih >= H) {
            std::fill(dst + (size_t)h * W, dst + (size_t)(h + 1) * W, T(0));
            continue;
          }
          const T* src = x + ((size_t)c * H + ih) * W;
          for (int w = 0; w < W; ++w) {
            const int iw = w + kw - pad;
            dst[(size_t)h * W + w] = (iw < 0 || iw >= W) ? T(0) : src[iw];
          }
        }
      }
}
template <class T>
void col2im_add(const T* col, int Cin, int H, int W, int k, T* dx) {
  const int pad = (k - 1) / 2;
  for (int c = 0; c < Cin; ++c)
    for (int kh = 0; kh < k; ++kh)
      for (int kw = 0; kw < k; ++kw) {
        const T* src = col + ((size_t)(c * k + kh) * k + kw) * H * W;
        for (int h = 0; h < H; ++h) {
          const int ih = h + kh - pad;
          if (ih < 0 || ih >= H) continue;
          T* dst = dx + ((size_t)c * H + ih) * W;
          for (int w = 0; w < W; ++w) {
            const int iw = w + kw - pad;
            if (iw >= 0 && iw < W) dst[iw] += src[(size_t)h * W + w];
          }
        }
      }
}
// im2col of ONE sample, rows (c,kh,kw) distributed over the threads of the enclosing parallel region
template <class T>
void im2col_rows(const T* x, int Cin, int H, int W, int k, T* col) {
  const int pad = (k - 1) / 2, KK = k * k;
#pragma omp for schedule(static)
  for (int r = 0; r < Cin * KK; ++r) {
    const int c = r / KK, kh = (r % KK) / k, kw = r % k;
    T* dst = col + (size_t)r * H * W;
    for (int h = 0; h < H; ++h) {
      const int ih = h + kh - pad;
      if (ih < 0 || ih >= H) {
        std::fill(dst + (size_t)h * W, dst + (size_t)(h + 1) * W, T(0));
        continue;
      }
      const T* src = x + ((size_t)c * H + ih) * W;
      for (int w = 0; w < W; ++w) {
        const int iw = w + kw - pad;
        dst[(size_t)h * W + w] = (iw < 0 || iw >= W) ? T(0) : src[iw];
      }
    }
  }
}
// THNN SpatialConvolutionMM: loop over samples, im2col + GEMM per sample; the threads share one sample's
// GEMM (rows of the output split across threads), so parallelism does not depend on the batch size.
// im2col / col2im of one sample with their own parallel region (BLAS path: the GEMM is threaded by the BLAS)
inline void im2col_par(const float* x, int Cin, int H, int W, int k, float* col) {
#pragma omp parallel
  im2col_rows(x, Cin, H, W, k, col);
}
inline void col2im_par(const float* col, int Cin, int H, int W, int k, float* dx) {
  const int pad = (k - 1) / 2, KK = k * k, HW = H * W;
#pragma omp parallel for schedule(static)
  for (int c = 0; c < Cin; ++c) {
    float* dst_c = dx + (size_t)c * HW;
    std::fill(dst_c, dst_c + HW, 0.f);
    for (int t = 0; t < KK; ++t) {
      const int kh = t / k, kw = t % k;
      const float* src = col + ((size_t)c * KK + t) * HW;
      for (int h = 0; h < H; ++h) {
        const int ih = h + kh - pad;
        if (ih < 0 || ih >= H) continue;
        for (int w = 0; w < W; ++w) {
          const int iw = w + kw - pad;
          if (iw >= 0 && iw < W) dst_c[(size_t)ih * W + iw] += src[(size_t)h * W + w];
        }
      }
    }
  }
}

template <class T>
void conv_fwd(int B, int Cin, int H, int W, int Cout, int k, const T* x, const T* Wt, const T* b, T* y) {
  const int K = Cin * k * k, HW = H * W;
  std::vector<T> col((size_t)K * HW);
  if constexpr (std::is_same<T, float>::value) {
    if (g_sgemm) {  // THNN SpatialConvolutionMM_updateOutput_frame: output = weight * finput (+ bias)
      for (int n = 0; n < B; ++n) {
        im2col_par(x + (size_t)n * Cin * HW, Cin, H, W, k, col.data());
        float* yn = y + (size_t)n * Cout * HW;
        g_sgemm(kRowMajor, kNoTrans, kNoTrans, Cout, HW, K, 1.f, Wt, K, col.data(), HW, 0.f, yn, HW);
        for (int o = 0; o < Cout; ++o)
          for (int p = 0; p < HW; ++p) yn[(size_t)o * HW + p] += b[o];
      }
      return;
    }
  }
#pragma omp parallel
  for (int n = 0; n < B; ++n) {
    im2col_rows(x + (size_t)n * Cin * HW, Cin, H, W, k, col.data());  // implicit barrier after the omp for
    T* yn = y + (size_t)n * Cout * HW;
#pragma omp for schedule(static)
    for (int ob = 0; ob < (Cout + 3) / 4; ++ob) {
      const int o0 = ob * 4, no = std::min(4, Cout - o0);
      gemm_nn(no, HW, K, Wt + (size_t)o0 * K, col.data(), yn + (size_t)o0 * HW, false);
      for (int o = o0; o < o0 + no; ++o)
        for (int p = 0; p < HW; ++p) yn[(size_t)o * HW + p] += b[o];
    }
  }
}
// dx (may be null) = conv^T(dy) ; dW += ... ; db += ...
template <class T>
void conv_bwd(int B, int Cin, int H, int W, int Cout, int k, const T* x, const T* Wt, const T* dy, T* dx,
              T* dW, T* db) {
  const int K = Cin * k * k, HW = H * W, KK = k * k, pad = (k - 1) / 2;
  std::vector<T> col((size_t)K * HW);
  if constexpr (std::is_same<T, float>::value) {
    if (g_sgemm) {  // THNN updateGradInput_frame: fgradInput = weight^T * gradOutput; accGradParameters: gradWeight += gradOutput * finput^T
      for (int n = 0; n < B; ++n) {
        const float* dyn = dy + (size_t)n * Cout * HW;
        if (dx) {
          g_sgemm(kRowMajor, kTrans, kNoTrans, K, HW, Cout, 1.f, Wt, K, dyn, HW, 0.f, col.data(), HW);
          col2im_par(col.data(), Cin, H, W, k, dx + (size_t)n * Cin * HW);
        }
        if (dW) {
          im2col_par(x + (size_t)n * Cin * HW, Cin, H, W, k, col.data());
          g_sgemm(kRowMajor, kNoTrans, kTrans, Cout, K, HW, 1.f, dyn, HW, col.data(), HW, 1.f, dW, K);
        }
      }
      if (db) {
#pragma omp parallel for schedule(static)
        for (int o = 0; o < Cout; ++o) {
          float s = 0;
          for (int n = 0; n < B; ++n)
            for (int p = 0; p < HW; ++p) s += dy[((size_t)n * Cout + o) * HW + p];
          db[o] += s;
        }
      }
      return;
    }
  }
  std::vector<T> WtT;
  if (dx) {  // W^T [K][Cout] so that column-gradient rows are independent dot products
    WtT.resize((size_t)K * Cout);
    for (int o = 0; o < Cout; ++o)
      for (int r = 0; r < K; ++r) WtT[(size_t)r * Cout + o] = Wt[(size_t)o * K + r];
  }
#pragma omp parallel
  for (int n = 0; n < B; ++n) {
    const T* dyn = dy + (size_t)n * Cout * HW;
    if (dx) {
#pragma omp for schedule(static)
      for (int rb = 0; rb < (K + 3) / 4; ++rb) {
        const int r0 = rb * 4, nr = std::min(4, K - r0);
        gemm_nn(nr, HW, Cout, WtT.data() + (size_t)r0 * Cout, dyn, col.data() + (size_t)r0 * HW, false);
      }
      T* dxn = dx + (size_t)n * Cin * HW;
#pragma omp for schedule(static)
      for (int c = 0; c < Cin; ++c) {  // col2im: each thread owns whole input channels
        T* dst_c = dxn + (size_t)c * HW;
        std::fill(dst_c, dst_c + HW, T(0));
        for (int t = 0; t < KK; ++t) {
          const int kh = t / k, kw = t % k;
          const T* src = col.data() + ((size_t)c * KK + t) * HW;
          for (int h = 0; h < H; ++h) {
            const int ih = h + kh - pad;
            if (ih < 0 || ih >= H) continue;
            for (int w = 0; w < W; ++w) {
              const int iw = w + kw - pad;
              if (iw >= 0 && iw < W) dst_c[(size_t)ih * W + iw] += src[(size_t)h * W + w];
            }
          }
        }
      }
    }
    if (dW) {
      im2col_rows(x + (size_t)n * Cin * HW, Cin, H, W, k, col.data());
#pragma omp for schedule(static)
      for (int ob = 0; ob < (Cout + 3) / 4; ++ob) {
        const int o0 = ob * 4, no = std::min(4, Cout - o0);
        gemm_nt(no, K, HW, dyn + (size_t)o0 * HW, col.data(), dW + (size_t)o0 * K, true);
      }
    }
  }
  if (db) {
#pragma omp parallel for schedule(static)
    for (int o = 0; o < Cout; ++o) {
      T s = 0;
      for (int n = 0; n < B; ++n)
        for (int p = 0; p < HW; ++p) s += dy[((size_t)n * Cout + o) * HW + p];
      db[o] += s;
    }
  }
}

// nn.SpatialUpSamplingNearest(2)                                    (models.lua:63,68)
template <class T>
void up2_fwd(int B, int C, int H, int W, const T* x, T* y) {
#pragma omp parallel for schedule(static)
  for (int nc = 0; nc < B * C; ++nc)
    for (int h = 0; h < 2 * H; ++h)
      for (int w = 0; w < 2 * W; ++w)
        y[((size_t)nc * 2 * H + h) * 2 * W + w] = x[((size_t)nc * H + h / 2) * W + w / 2];
}
template <class T>
void up2_bwd(int B, int C, int H, int W, const T* dy, T* dx) {
#pragma omp parallel for schedule(static)
  for (int nc = 0; nc < B * C; ++nc)
    for (int h = 0; h < H; ++h)
      for (int w = 0; w < W; ++w) {
        const T* r0 = dy + ((size_t)nc * 2 * H + 2 * h) * 2 * W + 2 * w;
        const T* r1 = r0 + 2 * W;
        dx[((size_t)nc * H + h) * W + w] = r0[0] + r0[1] + r1[0] + r1[1];
      }
}

// nn.SpatialBatchNormalization(C): eps 1e-5, momentum 0.1, affine, training mode (models.lua:65,70)
template <class T>
struct BNSave {
  std::vector<T> mean, istd;
};
template <class T>
void bn_fwd_train(int B, int C, int HW, const T* x, const T* gamma, const T* beta, T* y, BNSave<T>& s,
                  T* run_mean, T* run_var) {
  const T eps = T(1e-5), mom = T(0.1);
  s.mean.assign(C, 0);
  s.istd.assign(C, 0);
  const size_t cnt = (size_t)B * HW;
#pragma omp parallel for schedule(static)
  for (int c = 0; c < C; ++c) {
    T m = 0;
    for (int n = 0; n < B; ++n)
      for (int p = 0; p < HW; ++p) m += x[((size_t)n * C + c) * HW + p];
    m /= T(cnt);
    T v = 0;
    for (int n = 0; n < B; ++n)
      for (int p = 0; p < HW; ++p) {
        const T d = x[((size_t)n * C + c) * HW + p] - m;
        v += d * d;
      }
    const T var_b = v / T(cnt);
    const T istd = T(1) / std::sqrt(var_b + eps);
    s.mean[c] = m;
    s.istd[c] = istd;
    if (run_mean) run_mean[c] = (T(1) - mom) * run_mean[c] + mom * m;
    if (run_var) run_var[c] = (T(1) - mom) * run_var[c] + mom * (cnt > 1 ? v / T(cnt - 1) : var_b);
    for (int n = 0; n < B; ++n)
      for (int p = 0; p < HW; ++p) {
        const size_t i = ((size_t)n * C + c) * HW + p;
        y[i] = gamma[c] * ((x[i] - m) * istd) + beta[c];
      }
  }
}
template <class T>
void bn_fwd_eval(int B, int C, int HW, const T* x, const T* gamma, const T* beta, T* y, const T* run_mean,
                 const T* run_var) {
  const T eps = T(1e-5);
  for (int c = 0; c < C; ++c) {
    const T istd = T(1) / std::sqrt(run_var[c] + eps);
    for (int n = 0; n < B; ++n)
      for (int p = 0; p < HW; ++p) {
        const size_t i = ((size_t)n * C + c) * HW + p;
        y[i] = gamma[c] * ((x[i] - run_mean[c]) * istd) + beta[c];
      }
  }
}
// dx = gamma*istd*(g - mean(g) - xhat*mean(g*xhat)); dgamma += sum g*xhat; dbeta += sum g
template <class T>
void bn_bwd(int B, int C, int HW, const T* x, const T* gamma, const BNSave<T>& s, const T* dy, T* dx, T* dgamma,
            T* dbeta) {
  const size_t cnt = (size_t)B * HW;
#pragma omp parallel for schedule(static)
  for (int c = 0; c < C; ++c) {
    T sg = 0, sgx = 0;
    for (int n = 0; n < B; ++n)
      for (int p = 0; p < HW; ++p) {
        const size_t i = ((size_t)n * C + c) * HW + p;
        const T xh = (x[i] - s.mean[c]) * s.istd[c];
        sg += dy[i];
        sgx += dy[i] * xh;
      }
    dgamma[c] += sgx;
    dbeta[c] += sg;
    const T mg = sg / T(cnt), mgx = sgx / T(cnt);
    for (int n = 0; n < B; ++n)
      for (int p = 0; p < HW; ++p) {
        const size_t i = ((size_t)n * C + c) * HW + p;
        const T xh = (x[i] - s.mean[c]) * s.istd[c];
        dx[i] = gamma[c] * s.istd[c] * (dy[i] - mg - xh * mgx);
      }
  }
}

// ---- PReLU kink bookkeeping (used by the parity tests only; off by default) ----------------------
// PReLU's derivative is discontinuous at 0.  A pre-activation that lies within rounding noise of 0 may be
// rounded to the other sign by a second (fp32) implementation; both branch choices are then "correct", but
// the two gradients differ by (1-a)*dy for that element.  To compare gradients STRICTLY the tests
//   1. run the oracle in RECORD mode: every prelu_fwd call (numbered in call order) lists the elements with
//      |x| < margin * max|x|  ("ambiguous" elements),
//   2. read the branch the implementation under test took for exactly those elements,
//   3. re-run the oracle in OVERRIDE mode: prelu_bwd takes the given branch for the listed elements and its
//      own branch everywhere else.
// The forward value is continuous in x, so only the backward is touched.  prelu_bwd finds its forward call
// through the address of the pre-activation array (the nets keep those arrays alive between the two).
struct KinkCall {
  size_t n = 0;
  double maxabs = 0;
  std::vector<long> idx;           // RECORD: ambiguous elements
  std::vector<long> ov_idx;        // OVERRIDE: elements whose branch is forced ...
  std::vector<signed char> ov_pos;  // ... to x > 0 (1) or x <= 0 (0)
};
struct KinkCtx {
  int mode = 0;  // 0 off, 1 record, 2 override
  double margin = 0;
  int next_call = 0;
  std::vector<KinkCall> calls;
  std::vector<std::pair<const void*, int>> ptr_call;  // latest forward call per pre-activation array
} g_kink;

template <class T>
void kink_on_forward(size_t n, const T* x) {
  const int s = g_kink.next_call++;
  if ((int)g_kink.calls.size() <= s) g_kink.calls.resize(s + 1);
  bool found = false;
  for (auto& pc : g_kink.ptr_call)
    if (pc.first == (const void*)x) { pc.second = s; found = true; }
  if (!found) g_kink.ptr_call.emplace_back((const void*)x, s);
  KinkCall& kc = g_kink.calls[s];
  kc.n = n;
  if (g_kink.mode == 1) {
    double mx = 0;
    for (size_t i = 0; i < n; ++i) mx = std::max(mx, (double)std::fabs(x[i]));
    kc.maxabs = mx;
    kc.idx.clear();
    const double thr = g_kink.margin * mx;
    for (size_t i = 0; i < n; ++i)
      if ((double)std::fabs(x[i]) < thr) kc.idx.push_back((long)i);
  }
}

// nn.PReLU() with nOutputPlane=0 => ONE shared slope         (models.lua:61,66,71,386,...)
template <class T>
void prelu_fwd(size_t n, const T* x, T a, T* y) {
  if (g_kink.mode) kink_on_forward(n, x);
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n; ++i) y[i] = x[i] > 0 ? x[i] : a * x[i];
}
template <class T>
void prelu_bwd(size_t n, const T* x, T a, const T* dy, T* dx, T* da) {
  T s = 0;
#pragma omp parallel for schedule(static) reduction(+ : s)
  for (size_t i = 0; i < n; ++i) {
    if (x[i] > 0) {
      dx[i] = dy[i];
    } else {
      dx[i] = a * dy[i];
      s += dy[i] * x[i];
    }
  }
  if (g_kink.mode == 2) {  // forced branches of the ambiguous elements (see above)
    int call = -1;
    for (const auto& pc : g_kink.ptr_call)
      if (pc.first == (const void*)x) call = pc.second;
    if (call >= 0 && call < (int)g_kink.calls.size()) {
      const KinkCall& kc = g_kink.calls[call];
      for (size_t j = 0; j < kc.ov_idx.size(); ++j) {
        const size_t i = (size_t)kc.ov_idx[j];
        if (i >= n) continue;
        const bool own = x[i] > 0, want = kc.ov_pos[j] != 0;
        if (own == want) continue;
        if (want) {  // take the x > 0 branch: undo the slope-gradient term, pass dy through
          s -= dy[i] * x[i];
          dx[i] = dy[i];
        } else {
          s += dy[i] * x[i];
          dx[i] = a * dy[i];
        }
      }
    }
  }
  *da += s;
}

// nn.SpatialAveragePooling(2,2,2,2)
template <class T>
void avgpool2_fwd(int BC, int H, int W, const T* x, T* y) {
  const int Ho = H / 2, Wo = W / 2;
#pragma omp parallel for schedule(static)
  for (int nc = 0; nc < BC; ++nc)
    for (int h = 0; h < Ho; ++h)
      for (int w = 0; w < Wo; ++w) {
        const T* r0 = x + ((size_t)nc * H + 2 * h) * W + 2 * w;
        y[((size_t)nc * Ho + h) * Wo + w] = (r0[0] + r0[1] + r0[W] + r0[W + 1]) * T(0.25);
      }
}
template <class T>
void avgpool2_bwd(int BC, int H, int W, const T* dy, T* dx) {
  const int Ho = H / 2, Wo = W / 2;
#pragma omp parallel for schedule(static)
  for (int nc = 0; nc < BC; ++nc)
    for (int h = 0; h < H; ++h)
      for (int w = 0; w < W; ++w)
        dx[((size_t)nc * H + h) * W + w] = dy[((size_t)nc * Ho + h / 2) * Wo + w / 2] * T(0.25);
}

// D's sigmoid output as the reference's criterion and backward pass see it: a FLOAT32 value.  The reference
// computes D on the GPU in fp32 and hands `outputs` to nn.BCECriterion through nn.Copy as a FloatTensor
// (train.lua:148; utils/nn_utils.lua:359), so a saturated output is EXACTLY 0 or 1: BCE's log(1 - x + eps) then sees
// eps alone and the composed gradient BCE' * y(1 - y) is exactly 0.  The fp64 oracle models
// that storage precision here (and only here); everything else stays double.
template <class T>
inline T d_output(T y) {
  return (T)(float)y;
}
template <class T>
inline T sigmoid(T x) {
  return T(1) / (T(1) + std::exp(-x));
}

// nn.BCECriterion (2015 Lua implementation; eps = 1e-12; sizeAverage)      (train.lua:148)
//   loss = -(1/N) sum[ t log(x+eps) + (1-t) log(1-x+eps) ]
//   grad = -(t-x) / ( x (1-x+eps) + eps ) / N
template <class T>
T bce_fwd(int N, const T* x, const T* t) {
  const T eps = T(1e-12);
  T s = 0;
  for (int i = 0; i < N; ++i) s += t[i] * std::log(x[i] + eps) + (T(1) - t[i]) * std::log(T(1) - x[i] + eps);
  return -s / T(N);
}
template <class T>
void bce_bwd(int N, const T* x, const T* t, T* dx) {
  const T eps = T(1e-12);
  for (int i = 0; i < N; ++i) dx[i] = -(t[i] - x[i]) / (x[i] * (T(1) - x[i] + eps) + eps) / T(N);
}

// ----------------------------------------------------------------------------------------------
// Flat parameter layouts (getParameters order)
// ----------------------------------------------------------------------------------------------
struct GLayout {
  int C;
  size_t L1W, L1b, a1, C1W, C1b, g1, be1, a2, C2W, C2b, g2, be2, a3, C3W, C3b, total;
  explicit GLayout(int C_) : C(C_) {
    size_t o = 0;
    L1W = o; o += 8192 * 100;
    L1b = o; o += 8192;
    a1 = o; o += 1;
    C1W = o; o += 256 * 128 * 25;
    C1b = o; o += 256;
    g1 = o; o += 256;
    be1 = o; o += 256;
    a2 = o; o += 1;
    C2W = o; o += 128 * 256 * 25;
    C2b = o; o += 128;
    g2 = o; o += 128;
    be2 = o; o += 128;
    a3 = o; o += 1;
    C3W = o; o += (size_t)C * 128 * 9;
    C3b = o; o += C;
    total = o;
  }
};
struct DLayout {
  int C;
  size_t cW[4], cb[4], ca[4], L1W, L1b, a5, L2W, L2b, a6, L3W, L3b, total;
  explicit DLayout(int C_) : C(C_) {
    const int cin[4] = {C_, 64, 128, 256}, cout[4] = {64, 128, 256, 512};
    size_t o = 0;
    for (int i = 0; i < 4; ++i) {
      cW[i] = o; o += (size_t)cout[i] * cin[i] * 9;
      cb[i] = o; o += cout[i];
      ca[i] = o; o += 1;
    }
    L1W = o; o += 512 * 2048;
    L1b = o; o += 512;
    a5 = o; o += 1;
    L2W = o; o += 512 * 512;
    L2b = o; o += 512;
    a6 = o; o += 1;
    L3W = o; o += 512;
    L3b = o; o += 1;
    total = o;
  }
};

// ----------------------------------------------------------------------------------------------
// Generator  (models.lua:57-81)
// ----------------------------------------------------------------------------------------------
template <class T>
struct GNet {
  int B = 0, C = 3;
  std::vector<T> x, z0, h0, u0, z1, y1, h1, u1, z2, y2, h2, z3, out;
  BNSave<T> s1, s2;
  void forward(const T* P, const T* noise, int B_, int C_, bool training, T* bn_state /*768: rm1,rv1,rm2,rv2*/) {
    B = B_; C = C_;
    GLayout L(C);
    x.assign(noise, noise + (size_t)B * 100);
    z0.resize((size_t)B * 8192); h0.resize(z0.size());
    linear_fwd(B, 100, 8192, x.data(), P + L.L1W, P + L.L1b, z0.data());
    prelu_fwd(z0.size(), z0.data(), P[L.a1], h0.data());
    u0.resize((size_t)B * 128 * 256);
    up2_fwd(B, 128, 8, 8, h0.data(), u0.data());
    z1.resize((size_t)B * 256 * 256); y1.resize(z1.size()); h1.resize(z1.size());
    conv_fwd(B, 128, 16, 16, 256, 5, u0.data(), P + L.C1W, P + L.C1b, z1.data());
    T* rm1 = bn_state; T* rv1 = bn_state ? bn_state + 256 : nullptr;
    T* rm2 = bn_state ? bn_state + 512 : nullptr; T* rv2 = bn_state ? bn_state + 640 : nullptr;
    if (training) bn_fwd_train(B, 256, 256, z1.data(), P + L.g1, P + L.be1, y1.data(), s1, rm1, rv1);
    else bn_fwd_eval(B, 256, 256, z1.data(), P + L.g1, P + L.be1, y1.data(), rm1, rv1);
    prelu_fwd(y1.size(), y1.data(), P[L.a2], h1.data());
    u1.resize((size_t)B * 256 * 1024);
    up2_fwd(B, 256, 16, 16, h1.data(), u1.data());
    z2.resize((size_t)B * 128 * 1024); y2.resize(z2.size()); h2.resize(z2.size());
    conv_fwd(B, 256, 32, 32, 128, 5, u1.data(), P + L.C2W, P + L.C2b, z2.data());
    if (training) bn_fwd_train(B, 128, 1024, z2.data(), P + L.g2, P + L.be2, y2.data(), s2, rm2, rv2);
    else bn_fwd_eval(B, 128, 1024, z2.data(), P + L.g2, P + L.be2, y2.data(), rm2, rv2);
    prelu_fwd(y2.size(), y2.data(), P[L.a3], h2.data());
    z3.resize((size_t)B * C * 1024); out.resize(z3.size());
    conv_fwd(B, 128, 32, 32, C, 3, h2.data(), P + L.C3W, P + L.C3b, z3.data());
    for (size_t i = 0; i < z3.size(); ++i) out[i] = sigmoid(z3[i]);
  }
  // grads ACCUMULATE into dP (caller zeroes, adversarial.lua:193). dnoise optional.
  void backward(const T* P, const T* dout, T* dP, T* dnoise) {
    GLayout L(C);
    std::vector<T> dz3(z3.size());
    for (size_t i = 0; i < dz3.size(); ++i) dz3[i] = dout[i] * out[i] * (T(1) - out[i]);
    std::vector<T> dh2(h2.size());
    conv_bwd(B, 128, 32, 32, C, 3, h2.data(), P + L.C3W, dz3.data(), dh2.data(), dP + L.C3W, dP + L.C3b);
    std::vector<T> dy2(y2.size()), dz2(z2.size());
    prelu_bwd(y2.size(), y2.data(), P[L.a3], dh2.data(), dy2.data(), dP + L.a3);
    bn_bwd(B, 128, 1024, z2.data(), P + L.g2, s2, dy2.data(), dz2.data(), dP + L.g2, dP + L.be2);
    std::vector<T> du1(u1.size()), dh1(h1.size());
    conv_bwd(B, 256, 32, 32, 128, 5, u1.data(), P + L.C2W, dz2.data(), du1.data(), dP + L.C2W, dP + L.C2b);
    up2_bwd(B, 256, 16, 16, du1.data(), dh1.data());
    std::vector<T> dy1(y1.size()), dz1(z1.size());
    prelu_bwd(y1.size(), y1.data(), P[L.a2], dh1.data(), dy1.data(), dP + L.a2);
    bn_bwd(B, 256, 256, z1.data(), P + L.g1, s1, dy1.data(), dz1.data(), dP + L.g1, dP + L.be1);
    std::vector<T> du0(u0.size()), dh0(h0.size());
    conv_bwd(B, 128, 16, 16, 256, 5, u0.data(), P + L.C1W, dz1.data(), du0.data(), dP + L.C1W, dP + L.C1b);
    up2_bwd(B, 128, 8, 8, du0.data(), dh0.data());
    std::vector<T> dz0(z0.size());
    prelu_bwd(z0.size(), z0.data(), P[L.a1], dh0.data(), dz0.data(), dP + L.a1);
    std::vector<T> dx;
    if (dnoise) dx.resize((size_t)B * 100);
    linear_bwd(B, 100, 8192, x.data(), P + L.L1W, dz0.data(), dnoise ? dx.data() : nullptr, dP + L.L1W, dP + L.L1b);
    if (dnoise) std::copy(dx.begin(), dx.end(), dnoise);
  }
};

// ----------------------------------------------------------------------------------------------
// Discriminator (models.lua:382-416).  Dropout masks are INPUTS (keep flags 0/1):
//   per sample: [64 | 128 | 256 | 512] SpatialDropout(0.2) channel masks (no rescale in training;
//   eval multiplies by 1-p), then [512 | 512] nn.Dropout(0.5) masks (v2: kept values / (1-p)).
// ----------------------------------------------------------------------------------------------
constexpr int kMaskPerSample = 64 + 128 + 256 + 512 + 512 + 512;  // 1984
template <class T>
struct DNet {
  int B = 0, C = 3;
  std::vector<T> x;
  std::vector<T> z[4], a[4], d[4], p[4];  // conv out, prelu out, dropout out, pooled
  std::vector<T> zl1, al1, hl1, zl2, al2, hl2, logit, out;
  std::vector<T> mask;
  bool training = true;
  void forward(const T* P, const T* img, int B_, int C_, bool training_, const T* masks) {
    B = B_; C = C_; training = training_;
    DLayout L(C);
    const int cin[4] = {C, 64, 128, 256}, cout[4] = {64, 128, 256, 512}, hw[4] = {32, 16, 8, 4};
    const int moff[4] = {0, 64, 192, 448};
    x.assign(img, img + (size_t)B * C * 1024);
    if (training) mask.assign(masks, masks + (size_t)B * kMaskPerSample);
    const T* cur = x.data();
    for (int i = 0; i < 4; ++i) {
      const int H = hw[i];
      const size_t n = (size_t)B * cout[i] * H * H;
      z[i].resize(n); a[i].resize(n); d[i].resize(n); p[i].resize(n / 4);
      conv_fwd(B, cin[i], H, H, cout[i], 3, cur, P + L.cW[i], P + L.cb[i], z[i].data());
      prelu_fwd(n, z[i].data(), P[L.ca[i]], a[i].data());
      for (int b = 0; b < B; ++b)
        for (int c = 0; c < cout[i]; ++c) {
          const T m = training ? mask[(size_t)b * kMaskPerSample + moff[i] + c] : T(0.8);
          for (int q = 0; q < H * H; ++q) {
            const size_t idx = ((size_t)b * cout[i] + c) * H * H + q;
            d[i][idx] = a[i][idx] * m;
          }
        }
      avgpool2_fwd(B * cout[i], H, H, d[i].data(), p[i].data());
      cur = p[i].data();
    }
    // View(2048): flatten C,H,W order == NCHW memory order of p[3] ([B][512][2][2])
    zl1.resize((size_t)B * 512); al1.resize(zl1.size()); hl1.resize(zl1.size());
    linear_fwd(B, 2048, 512, p[3].data(), P + L.L1W, P + L.L1b, zl1.data());
    prelu_fwd(zl1.size(), zl1.data(), P[L.a5], al1.data());
    for (int b = 0; b < B; ++b)
      for (int j = 0; j < 512; ++j)
        hl1[(size_t)b * 512 + j] = training ? al1[(size_t)b * 512 + j] * mask[(size_t)b * kMaskPerSample + 960 + j] * T(2)
                                           : al1[(size_t)b * 512 + j];
    zl2.resize(zl1.size()); al2.resize(zl1.size()); hl2.resize(zl1.size());
    linear_fwd(B, 512, 512, hl1.data(), P + L.L2W, P + L.L2b, zl2.data());
    prelu_fwd(zl2.size(), zl2.data(), P[L.a6], al2.data());
    for (int b = 0; b < B; ++b)
      for (int j = 0; j < 512; ++j)
        hl2[(size_t)b * 512 + j] = training ? al2[(size_t)b * 512 + j] * mask[(size_t)b * kMaskPerSample + 1472 + j] * T(2)
                                           : al2[(size_t)b * 512 + j];
    logit.resize(B); out.resize(B);
    linear_fwd(B, 512, 1, hl2.data(), P + L.L3W, P + L.L3b, logit.data());
    for (int b = 0; b < B; ++b) out[b] = d_output(sigmoid(logit[b]));
  }
  // dout: [B] gradient wrt sigmoid output. grads accumulate into dP (may be null => skip
  // weight grads). dimg optional [B][C][32][32].
  void backward(const T* P, const T* dout, T* dP, T* dimg) {
    DLayout L(C);
    const int cin[4] = {C, 64, 128, 256}, cout[4] = {64, 128, 256, 512}, hw[4] = {32, 16, 8, 4};
    const int moff[4] = {0, 64, 192, 448};
    std::vector<T> scratch(L.total, T(0));
    T* G = dP ? dP : scratch.data();
    std::vector<T> dlogit(B);
    for (int b = 0; b < B; ++b) dlogit[b] = dout[b] * out[b] * (T(1) - out[b]);
    std::vector<T> dhl2((size_t)B * 512), dal2(dhl2.size()), dzl2(dhl2.size());
    linear_bwd(B, 512, 1, hl2.data(), P + L.L3W, dlogit.data(), dhl2.data(), G + L.L3W, G + L.L3b);
    for (int b = 0; b < B; ++b)
      for (int j = 0; j < 512; ++j)
        dal2[(size_t)b * 512 + j] =
            training ? dhl2[(size_t)b * 512 + j] * mask[(size_t)b * kMaskPerSample + 1472 + j] * T(2) : dhl2[(size_t)b * 512 + j];
    prelu_bwd(dzl2.size(), zl2.data(), P[L.a6], dal2.data(), dzl2.data(), G + L.a6);
    std::vector<T> dhl1(dhl2.size()), dal1(dhl2.size()), dzl1(dhl2.size());
    linear_bwd(B, 512, 512, hl1.data(), P + L.L2W, dzl2.data(), dhl1.data(), G + L.L2W, G + L.L2b);
    for (int b = 0; b < B; ++b)
      for (int j = 0; j < 512; ++j)
        dal1[(size_t)b * 512 + j] =
            training ? dhl1[(size_t)b * 512 + j] * mask[(size_t)b * kMaskPerSample + 960 + j] * T(2) : dhl1[(size_t)b * 512 + j];
    prelu_bwd(dzl1.size(), zl1.data(), P[L.a5], dal1.data(), dzl1.data(), G + L.a5);
    std::vector<T> dp((size_t)B * 2048);
    linear_bwd(B, 2048, 512, p[3].data(), P + L.L1W, dzl1.data(), dp.data(), G + L.L1W, G + L.L1b);
    for (int i = 3; i >= 0; --i) {
      const int H = hw[i];
      const size_t n = (size_t)B * cout[i] * H * H;
      std::vector<T> dd(n), da(n), dz(n);
      avgpool2_bwd(B * cout[i], H, H, dp.data(), dd.data());
      for (int b = 0; b < B; ++b)
        for (int c = 0; c < cout[i]; ++c) {
          const T m = training ? mask[(size_t)b * kMaskPerSample + moff[i] + c] : T(0.8);
          for (int q = 0; q < H * H; ++q) {
            const size_t idx = ((size_t)b * cout[i] + c) * H * H + q;
            da[idx] = dd[idx] * m;
          }
        }
      prelu_bwd(n, z[i].data(), P[L.ca[i]], da.data(), dz.data(), G + L.ca[i]);
      const T* in = i == 0 ? x.data() : p[i - 1].data();
      const bool need_dx = i > 0 || dimg != nullptr;
      std::vector<T> dx;
      if (need_dx) dx.resize((size_t)B * cin[i] * H * H);
      conv_bwd(B, cin[i], H, H, cout[i], 3, in, P + L.cW[i], dz.data(), need_dx ? dx.data() : nullptr, G + L.cW[i],
               G + L.cb[i]);
      if (i == 0) {
        if (dimg) std::copy(dx.begin(), dx.end(), dimg);
      } else {
        dp.swap(dx);
      }
    }
  }
};

// penalty (adversarial.lua:103-109 / :218-224) + clamp (:121-123 / :226-228); returns loss term
template <class T>
T penalty_clamp(size_t n, const T* p, T* g, T l1_loss, T l1_grad, T l2, T clampv) {
  T add = 0;
  if (l1_loss != 0 || l2 != 0) {
    T n1 = 0, n2 = 0;
    for (size_t i = 0; i < n; ++i) { n1 += std::fabs(p[i]); n2 += p[i] * p[i]; }
    add = l1_loss * n1 + l2 * n2 / T(2);
    for (size_t i = 0; i < n; ++i) {
      const T sg = p[i] > 0 ? T(1) : (p[i] < 0 ? T(-1) : T(0));
      g[i] += sg * l1_grad + p[i] * l2;
    }
  }
  if (clampv != 0)
    for (size_t i = 0; i < n; ++i) g[i] = std::min(std::max(g[i], -clampv), clampv);
  return add;
}

// interruptableAdam (interruptable_optimizers.lua:49-94); t is the value AFTER increment
template <class T>
void adam(size_t n, T* x, const T* g, T* m, T* v, int t, double lr, double b1, double b2, double eps) {
  const double bc1 = 1.0 - std::pow(b1, t), bc2 = 1.0 - std::pow(b2, t);
  const T step = T(lr * std::sqrt(bc2) / bc1);
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n; ++i) {
    m[i] = m[i] * T(b1) + T(1 - b1) * g[i];
    v[i] = v[i] * T(b2) + T(1 - b2) * g[i] * g[i];
    const T denom = std::sqrt(v[i]) + T(eps);
    x[i] -= step * m[i] / denom;
  }
}

struct Hyper {  // mirrors fg_hyper in include/fg_b200.h
  double lr_D, lr_G, beta1, beta2, eps;
  double D_L1, D_L2, G_L1, G_L2, D_clamp, G_clamp;
};

// One iteration of the adversarial.lua loop body (D_iterations = G_iterations = 1, no gating).
//   real[B/2,C,32,32], noiseD[B/2,100], noiseG[B,100], masksD[B,1984], masksG[B,1984]
//   state: PD,PG params (updated in place), mD,vD,mG,vG Adam moments, tD,tG counters, bnG[768]
//   outs: stats[8] = lossD, lossG, confusion counts (pred1&t1, pred0&t1, pred1&t0, pred0&t0), 0,0
//         gradD_out/gradG_out (post penalty+clamp, optional), fake_out[B/2,C,32,32] optional,
//         dOutD[B] D-step outputs optional
template <class T>
void train_iteration(int B, int C, const Hyper& hp, const T* real, const T* noiseD, const T* noiseG,
                     const T* masksD, const T* masksG, T* PD, T* PG, T* mD, T* vD, T* mG, T* vG, int* tD, int* tG,
                     T* bnG, double* stats, T* gradD_out, T* gradG_out, T* fake_out, T* outD_out) {
  GLayout LG(C);
  DLayout LD(C);
  const int Bh = B / 2;
  const size_t img = (size_t)C * 1024;
  GNet<T> G;
  DNet<T> D;
  // ---- D step (adversarial.lua:240-268) ----
  G.forward(PG, noiseD, Bh, C, true, bnG);  // createImages => G in training mode (nn_utils.lua:52)
  if (fake_out) std::copy(G.out.begin(), G.out.end(), fake_out);
  std::vector<T> inputs((size_t)B * img), targets(B);
  std::copy(real, real + Bh * img, inputs.begin());
  std::copy(G.out.begin(), G.out.end(), inputs.begin() + Bh * img);
  for (int i = 0; i < B; ++i) targets[i] = i < Bh ? T(1) : T(0);  // Y_NOT_GENERATOR=1, Y_GENERATOR=0
  std::vector<T> gD(LD.total, T(0));
  D.forward(PD, inputs.data(), B, C, true, masksD);
  if (outD_out) std::copy(D.out.begin(), D.out.end(), outD_out);
  T fD = bce_fwd(B, D.out.data(), targets.data());
  std::vector<T> df(B);
  bce_bwd(B, D.out.data(), targets.data(), df.data());
  D.backward(PD, df.data(), gD.data(), nullptr);
  fD += penalty_clamp(LD.total, PD, gD.data(), T(hp.D_L1), T(hp.D_L1), T(hp.D_L2), T(hp.D_clamp));
  double conf[4] = {0, 0, 0, 0};
  for (int i = 0; i < B; ++i) {
    const bool pred1 = D.out[i] > T(0.5);
    const bool t1 = i < Bh;
    conf[(pred1 ? 0 : 1) + (t1 ? 0 : 2)] += 1;
  }
  if (gradD_out) std::copy(gD.begin(), gD.end(), gradD_out);
  *tD += 1;
  adam(LD.total, PD, gD.data(), mD, vD, *tD, hp.lr_D, hp.beta1, hp.beta2, hp.eps);
  // ---- G step (adversarial.lua:275-288) ----
  std::vector<T> gG(LG.total, T(0));
  G.forward(PG, noiseG, B, C, true, bnG);
  for (int i = 0; i < B; ++i) targets[i] = T(1);
  D.forward(PD, G.out.data(), B, C, true, masksG);
  T fG = bce_fwd(B, D.out.data(), targets.data());
  bce_bwd(B, D.out.data(), targets.data(), df.data());
  std::vector<T> dimg((size_t)B * img);
  D.backward(PD, df.data(), nullptr, dimg.data());  // D's weight grads are discarded (zeroed at :92)
  G.backward(PG, dimg.data(), gG.data(), nullptr);
  // quirk: the L1 *gradient* term is scaled by G_L2 (adversarial.lua:223)
  fG += penalty_clamp(LG.total, PG, gG.data(), T(hp.G_L1), T(hp.G_L2), T(hp.G_L2), T(hp.G_clamp));
  if (gradG_out) std::copy(gG.begin(), gG.end(), gradG_out);
  *tG += 1;
  adam(LG.total, PG, gG.data(), mG, vG, *tG, hp.lr_G, hp.beta1, hp.beta2, hp.eps);
  stats[0] = (double)fD; stats[1] = (double)fG;
  stats[2] = conf[0]; stats[3] = conf[1]; stats[4] = conf[2]; stats[5] = conf[3];
  stats[6] = 0; stats[7] = 0;
}

#include "fg_oracle_c2f.h"  // coarse-to-fine nets + loop (models_c2f.lua, adversarial_c2f.lua)
#include "fg_oracle_s16.h"  // --scale 16 nets (models.lua:26-51, :279-316)

}  // namespace

// ==============================================================================================
// C exports (ctypes).  X(name, T) instantiates each entry for double (_f64) and float (_f32).
// ==============================================================================================
#define FG_EXPORTS(SFX, T)                                                                                       \
  extern "C" {                                                                                                   \
  void fgo_linear_fwd_##SFX(int B, int in, int out, const T* x, const T* W, const T* b, T* y) {                  \
    linear_fwd<T>(B, in, out, x, W, b, y);                                                                       \
  }                                                                                                              \
  void fgo_linear_bwd_##SFX(int B, int in, int out, const T* x, const T* W, const T* dy, T* dx, T* dW, T* db) {  \
    linear_bwd<T>(B, in, out, x, W, dy, dx, dW, db);                                                             \
  }                                                                                                              \
  void fgo_conv_fwd_##SFX(int B, int Cin, int H, int W, int Cout, int k, const T* x, const T* Wt, const T* b,    \
                          T* y) {                                                                                \
    conv_fwd<T>(B, Cin, H, W, Cout, k, x, Wt, b, y);                                                             \
  }                                                                                                              \
  void fgo_conv_bwd_##SFX(int B, int Cin, int H, int W, int Cout, int k, const T* x, const T* Wt, const T* dy,   \
                          T* dx, T* dW, T* db) {                                                                 \
    conv_bwd<T>(B, Cin, H, W, Cout, k, x, Wt, dy, dx, dW, db);                                                   \
  }                                                                                                              \
  void fgo_up2_fwd_##SFX(int B, int C, int H, int W, const T* x, T* y) { up2_fwd<T>(B, C, H, W, x, y); }         \
  void fgo_up2_bwd_##SFX(int B, int C, int H, int W, const T* dy, T* dx) { up2_bwd<T>(B, C, H, W, dy, dx); }     \
  void fgo_bn_fwd_train_##SFX(int B, int C, int HW, const T* x, const T* g, const T* be, T* y, T* mean,          \
                              T* istd, T* rm, T* rv) {                                                           \
    BNSave<T> s;                                                                                                 \
    bn_fwd_train<T>(B, C, HW, x, g, be, y, s, rm, rv);                                                           \
    std::copy(s.mean.begin(), s.mean.end(), mean);                                                               \
    std::copy(s.istd.begin(), s.istd.end(), istd);                                                               \
  }                                                                                                              \
  void fgo_bn_bwd_##SFX(int B, int C, int HW, const T* x, const T* g, const T* mean, const T* istd,              \
                        const T* dy, T* dx, T* dg, T* db) {                                                      \
    BNSave<T> s;                                                                                                 \
    s.mean.assign(mean, mean + C);                                                                               \
    s.istd.assign(istd, istd + C);                                                                               \
    bn_bwd<T>(B, C, HW, x, g, s, dy, dx, dg, db);                                                                \
  }                                                                                                              \
  void fgo_prelu_fwd_##SFX(long n, const T* x, T a, T* y) { prelu_fwd<T>((size_t)n, x, a, y); }                  \
  void fgo_prelu_bwd_##SFX(long n, const T* x, T a, const T* dy, T* dx, T* da) {                                 \
    prelu_bwd<T>((size_t)n, x, a, dy, dx, da);                                                                   \
  }                                                                                                              \
  void fgo_avgpool2_fwd_##SFX(int BC, int H, int W, const T* x, T* y) { avgpool2_fwd<T>(BC, H, W, x, y); }       \
  void fgo_avgpool2_bwd_##SFX(int BC, int H, int W, const T* dy, T* dx) { avgpool2_bwd<T>(BC, H, W, dy, dx); }   \
  double fgo_bce_fwd_##SFX(int N, const T* x, const T* t) { return (double)bce_fwd<T>(N, x, t); }                \
  void fgo_bce_bwd_##SFX(int N, const T* x, const T* t, T* dx) { bce_bwd<T>(N, x, t, dx); }                      \
  double fgo_penalty_clamp_##SFX(long n, const T* p, T* g, double l1_loss, double l1_grad, double l2,            \
                                 double clampv) {                                                                \
    return (double)penalty_clamp<T>((size_t)n, p, g, T(l1_loss), T(l1_grad), T(l2), T(clampv));                  \
  }                                                                                                              \
  void fgo_adam_##SFX(long n, T* x, const T* g, T* m, T* v, int t, double lr, double b1, double b2,              \
                      double eps) {                                                                              \
    adam<T>((size_t)n, x, g, m, v, t, lr, b1, b2, eps);                                                          \
  }                                                                                                              \
  /* whole nets: handle-based so forward state survives until backward */                                       \
  void* fgo_G_new_##SFX() { return new GNet<T>(); }                                                              \
  void fgo_G_free_##SFX(void* h) { delete (GNet<T>*)h; }                                                         \
  void fgo_G_forward_##SFX(void* h, const T* P, const T* noise, int B, int C, int training, T* bn_state,         \
                           T* out) {                                                                             \
    GNet<T>* g = (GNet<T>*)h;                                                                                    \
    g->forward(P, noise, B, C, training != 0, bn_state);                                                         \
    std::copy(g->out.begin(), g->out.end(), out);                                                                \
  }                                                                                                              \
  void fgo_G_backward_##SFX(void* h, const T* P, const T* dout, T* dP, T* dnoise) {                              \
    ((GNet<T>*)h)->backward(P, dout, dP, dnoise);                                                                \
  }                                                                                                              \
  /* intermediate taps for layer-level parity: which = 0:z0 1:h0 2:z1 3:h1 4:z2 5:h2 6:z3 */                     \
  long fgo_G_tap_##SFX(void* h, int which, T* dst) {                                                             \
    GNet<T>* g = (GNet<T>*)h;                                                                                    \
    std::vector<T>* v[] = {&g->z0, &g->h0, &g->z1, &g->h1, &g->z2, &g->h2, &g->z3};                              \
    if (which < 0 || which > 6) return -1;                                                                       \
    if (dst) std::copy(v[which]->begin(), v[which]->end(), dst);                                                 \
    return (long)v[which]->size();                                                                               \
  }                                                                                                              \
  void* fgo_D_new_##SFX() { return new DNet<T>(); }                                                              \
  void fgo_D_free_##SFX(void* h) { delete (DNet<T>*)h; }                                                         \
  void fgo_D_forward_##SFX(void* h, const T* P, const T* img, int B, int C, int training, const T* masks,        \
                           T* out) {                                                                             \
    DNet<T>* d = (DNet<T>*)h;                                                                                    \
    d->forward(P, img, B, C, training != 0, masks);                                                              \
    std::copy(d->out.begin(), d->out.end(), out);                                                                \
  }                                                                                                              \
  void fgo_D_backward_##SFX(void* h, const T* P, const T* dout, T* dP, T* dimg) {                                \
    ((DNet<T>*)h)->backward(P, dout, dP, dimg);                                                                  \
  }                                                                                                              \
  void fgo_train_iteration_##SFX(int B, int C, const double* hp11, const T* real, const T* noiseD,               \
                                 const T* noiseG, const T* masksD, const T* masksG, T* PD, T* PG, T* mD, T* vD,  \
                                 T* mG, T* vG, int* tD, int* tG, T* bnG, double* stats, T* gradD_out,            \
                                 T* gradG_out, T* fake_out, T* outD_out) {                                       \
    Hyper hp{hp11[0], hp11[1], hp11[2], hp11[3], hp11[4], hp11[5], hp11[6], hp11[7], hp11[8], hp11[9],           \
             hp11[10]};                                                                                          \
    train_iteration<T>(B, C, hp, real, noiseD, noiseG, masksD, masksG, PD, PG, mD, vD, mG, vG, tD, tG, bnG,      \
                       stats, gradD_out, gradG_out, fake_out, outD_out);                                         \
  }                                                                                                              \
  }

FG_EXPORTS(f64, double)
FG_EXPORTS(f32, float)

#define FG_C2F_EXPORTS(SFX, T)                                                                                   \
  extern "C" {                                                                                                   \
  void fgo_maxpool2_fwd_##SFX(int BC, int H, int W, const T* x, T* y, unsigned char* arg) {                      \
    maxpool2_fwd<T>(BC, H, W, x, y, arg);                                                                        \
  }                                                                                                              \
  void fgo_maxpool2_bwd_##SFX(int BC, int H, int W, const T* dy, const unsigned char* arg, T* dx) {              \
    maxpool2_bwd<T>(BC, H, W, dy, arg, dx);                                                                      \
  }                                                                                                              \
  void* fgo_c2f_G_new_##SFX() { return new C2fGNet<T>(); }                                                       \
  void fgo_c2f_G_free_##SFX(void* h) { delete (C2fGNet<T>*)h; }                                                  \
  void fgo_c2f_G_forward_##SFX(void* h, const T* P, const T* noise, const T* cond, int B, int C, T* out) {       \
    C2fGNet<T>* g = (C2fGNet<T>*)h;                                                                              \
    g->forward(P, noise, cond, B, C);                                                                            \
    std::copy(g->z[4].begin(), g->z[4].end(), out);                                                              \
  }                                                                                                              \
  void fgo_c2f_G_backward_##SFX(void* h, const T* P, const T* dout, T* dP) {                                     \
    ((C2fGNet<T>*)h)->backward(P, dout, dP);                                                                     \
  }                                                                                                              \
  void* fgo_c2f_D_new_##SFX() { return new C2fDNet<T>(); }                                                       \
  void fgo_c2f_D_free_##SFX(void* h) { delete (C2fDNet<T>*)h; }                                                  \
  void fgo_c2f_D_forward_##SFX(void* h, const T* P, const T* diff, const T* cond, int B, int C, int training,    \
                               const T* masks, T* out) {                                                         \
    C2fDNet<T>* d = (C2fDNet<T>*)h;                                                                              \
    d->forward(P, diff, cond, B, C, training != 0, masks);                                                       \
    std::copy(d->out.begin(), d->out.end(), out);                                                                \
  }                                                                                                              \
  void fgo_c2f_D_backward_##SFX(void* h, const T* P, const T* dout, T* dP, T* ddiff) {                           \
    ((C2fDNet<T>*)h)->backward(P, dout, dP, ddiff);                                                              \
  }                                                                                                              \
  void fgo_c2f_train_iteration_##SFX(int B, int C, const double* hp11, const T* real_diff, const T* condD,       \
                                     const T* noiseD, const T* condG, const T* noiseG, const T* masksD,          \
                                     const T* masksG, T* PD, T* PG, T* mD, T* vD, T* mG, T* vG, int* tD,         \
                                     int* tG, double* stats, T* gradD_out, T* gradG_out, T* fake_out,            \
                                     T* outD_out) {                                                              \
    Hyper hp{hp11[0], hp11[1], hp11[2], hp11[3], hp11[4], hp11[5], hp11[6], hp11[7], hp11[8], hp11[9],           \
             hp11[10]};                                                                                          \
    c2f_train_iteration<T>(B, C, hp, real_diff, condD, noiseD, condG, noiseG, masksD, masksG, PD, PG, mD, vD,    \
                           mG, vG, tD, tG, stats, gradD_out, gradG_out, fake_out, outD_out);                     \
  }                                                                                                              \
  }

FG_C2F_EXPORTS(f64, double)
FG_C2F_EXPORTS(f32, float)

#define FG_S16_EXPORTS(SFX, T)                                                                                   \
  extern "C" {                                                                                                   \
  void fgo_convs_fwd_##SFX(int B, int Cin, int H, int W, int Cout, int k, int stride, int pad, const T* x,       \
                           const T* Wt, const T* b, T* y) {                                                      \
    convs_fwd<T>(B, Cin, H, W, Cout, k, stride, pad, x, Wt, b, y);                                               \
  }                                                                                                              \
  void fgo_convs_bwd_##SFX(int B, int Cin, int H, int W, int Cout, int k, int stride, int pad, const T* x,       \
                           const T* Wt, const T* dy, T* dx, T* dW, T* db) {                                      \
    convs_bwd<T>(B, Cin, H, W, Cout, k, stride, pad, x, Wt, dy, dx, dW, db);                                     \
  }                                                                                                              \
  void* fgo_s16_G_new_##SFX() { return new G16Net<T>(); }                                                        \
  void fgo_s16_G_free_##SFX(void* h) { delete (G16Net<T>*)h; }                                                   \
  void fgo_s16_G_forward_##SFX(void* h, const T* P, const T* noise, int B, int C, T* bn_state, T* out) {         \
    G16Net<T>* g = (G16Net<T>*)h;                                                                                \
    g->forward(P, noise, B, C, bn_state);                                                                        \
    std::copy(g->out.begin(), g->out.end(), out);                                                                \
  }                                                                                                              \
  void fgo_s16_G_backward_##SFX(void* h, const T* P, const T* dout, T* dP) { ((G16Net<T>*)h)->backward(P, dout, dP); } \
  void* fgo_s16_D_new_##SFX() { return new D16Net<T>(); }                                                        \
  void fgo_s16_D_free_##SFX(void* h) { delete (D16Net<T>*)h; }                                                   \
  void fgo_s16_D_forward_##SFX(void* h, const T* P, const T* img, int B, int C, int training, const T* masks,    \
                               T* out) {                                                                         \
    D16Net<T>* d = (D16Net<T>*)h;                                                                                \
    d->forward(P, img, B, C, training != 0, masks);                                                              \
    std::copy(d->out.begin(), d->out.end(), out);                                                                \
  }                                                                                                              \
  void fgo_s16_D_backward_##SFX(void* h, const T* P, const T* dout, T* dP, T* dimg) {                            \
    ((D16Net<T>*)h)->backward(P, dout, dP, dimg);                                                                \
  }                                                                                                              \
  }

FG_S16_EXPORTS(f64, double)
FG_S16_EXPORTS(f32, float)

extern "C" {
long fgo_G_param_count(int C) { return (long)GLayout(C).total; }
// ---- PReLU kink bookkeeping for the parity tests (see KinkCtx) ----
// mode 1 = record ambiguous elements (|x| < margin * max|x|) per prelu_fwd call, 2 = apply overrides, 0 = off.
// Entering a mode restarts the call numbering; overrides survive until fgo_kink_clear().
void fgo_kink_mode(int mode, double margin) {
  g_kink.mode = mode;
  g_kink.margin = margin;
  g_kink.next_call = 0;
  g_kink.ptr_call.clear();
}
void fgo_kink_clear() { g_kink = KinkCtx(); }
int fgo_kink_num_calls() { return g_kink.next_call; }
long fgo_kink_call_info(int call, long* n, double* maxabs) {
  if (call < 0 || call >= (int)g_kink.calls.size()) return -1;
  if (n) *n = (long)g_kink.calls[call].n;
  if (maxabs) *maxabs = g_kink.calls[call].maxabs;
  return (long)g_kink.calls[call].idx.size();
}
void fgo_kink_call_indices(int call, long* dst) {
  if (call < 0 || call >= (int)g_kink.calls.size()) return;
  std::copy(g_kink.calls[call].idx.begin(), g_kink.calls[call].idx.end(), dst);
}
void fgo_kink_set_override(int call, long count, const long* idx, const signed char* positive) {
  if (call < 0) return;
  if ((int)g_kink.calls.size() <= call) g_kink.calls.resize(call + 1);
  g_kink.calls[call].ov_idx.assign(idx, idx + count);
  g_kink.calls[call].ov_pos.assign(positive, positive + count);
}

long fgo_D_param_count(int C) { return (long)DLayout(C).total; }
int fgo_mask_per_sample() { return kMaskPerSample; }
// Route the fp32 port's GEMMs through an OpenBLAS shared object (path = e.g. scipy.libs/libscipy_openblas-*.so);
// threads = BLAS threads (0 = leave as is).  Returns 1 on success, 0 if the library / symbol is not there.
int fgo_use_blas(const char* path, int threads) {
  g_sgemm = nullptr;
  if (!path || !*path) return 0;
  void* h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
  if (!h) return 0;
  void* f = dlsym(h, "scipy_cblas_sgemm");
  if (!f) f = dlsym(h, "cblas_sgemm");
  if (!f) return 0;
  typedef void (*setn_fn)(int);
  void* sn = dlsym(h, "scipy_openblas_set_num_threads");
  if (!sn) sn = dlsym(h, "openblas_set_num_threads");
  if (sn && threads > 0) ((setn_fn)sn)(threads);
  g_sgemm = (sgemm_fn)f;
  return 1;
}
int fgo_blas_active() { return g_sgemm != nullptr; }
long fgo_c2f_G_param_count(int C) { return (long)C2fGLayout(C).total; }
long fgo_c2f_D_param_count(int C) { return (long)C2fDLayout(C).total; }
int fgo_c2f_mask_per_sample() { return kC2fMaskPerSample; }
long fgo_s16_G_param_count(int C) { return (long)G16Layout(C).total; }
long fgo_s16_D_param_count(int C) { return (long)D16Layout(C).total; }
int fgo_s16_mask_per_sample() { return kD16MaskPerSample; }
int fgo_num_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void fgo_set_num_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}
}
