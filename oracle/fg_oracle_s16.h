// TEST INFRASTRUCTURE (see fg_oracle.cpp header): CPU restatement of the --scale 16 nets of train.lua.
// PARITY UNPINNED.  Included by fg_oracle.cpp inside its anonymous namespace.
//
//   G16 = models.lua:26-51   create_G_decoder_upsampling16: the 32x32 generator with every spatial size halved
//         (Linear(100, 128*4*4) -> View(128,4,4) ... -> C x 16 x 16)
//   D16 = models.lua:279-316 create_D16_d: ConcatTable{conv branch, dense branch} -> JoinTable(2) -> Linear(1152,1) -> Sigmoid
//         conv branch : conv(C->128,3,s1) PReLU conv(128->128,3,s1) PReLU AvgPool2 conv(128->512,3,STRIDE 2,pad 1) PReLU
//                       conv(512->1024,3,stride 2,pad 1) PReLU SpatialDropout() View(4096) Linear(4096,1024) PReLU
//         dense branch: View(C*256) Linear(C*256,128) PReLU Dropout Linear(128,128) PReLU
//   models.create_G / create_D pick these when dimensions[2] == 16 (models.lua:87-104).
// No CUDA counterpart exists yet; this restatement and its PyTorch cross-check
// (tests/test_oracle_s16_vs_torch.py) are the checker such an implementation will be held to.
#pragma once

// nn.SpatialConvolution(nIn, nOut, k, k, stride, stride, pad): cross-correlation, NCHW, Ho = (H + 2 pad - k)/stride + 1
template <class T>
void convs_fwd(int B, int Cin, int H, int W, int Cout, int k, int stride, int pad, const T* x, const T* Wt, const T* b, T* y) {
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
#pragma omp parallel for collapse(2) schedule(static)
  for (int n = 0; n < B; ++n)
    for (int o = 0; o < Cout; ++o) {
      T* yo = y + ((size_t)n * Cout + o) * Ho * Wo;
      for (int i = 0; i < Ho * Wo; ++i) yo[i] = b[o];
      for (int c = 0; c < Cin; ++c) {
        const T* xc = x + ((size_t)n * Cin + c) * H * W;
        const T* wk = Wt + ((size_t)o * Cin + c) * k * k;
        for (int kh = 0; kh < k; ++kh)
          for (int kw = 0; kw < k; ++kw) {
            const T w = wk[kh * k + kw];
            for (int ho = 0; ho < Ho; ++ho) {
              const int ih = ho * stride + kh - pad;
              if (ih < 0 || ih >= H) continue;
              for (int wo = 0; wo < Wo; ++wo) {
                const int iw = wo * stride + kw - pad;
                if (iw >= 0 && iw < W) yo[ho * Wo + wo] += w * xc[(size_t)ih * W + iw];
              }
            }
          }
      }
    }
}
// dx (may be null, overwritten); dW, db accumulate
template <class T>
void convs_bwd(int B, int Cin, int H, int W, int Cout, int k, int stride, int pad, const T* x, const T* Wt, const T* dy,
               T* dx, T* dW, T* db) {
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  if (dx) {
#pragma omp parallel for collapse(2) schedule(static)
    for (int n = 0; n < B; ++n)
      for (int c = 0; c < Cin; ++c) {
        T* dxc = dx + ((size_t)n * Cin + c) * H * W;
        std::fill(dxc, dxc + (size_t)H * W, T(0));
        for (int o = 0; o < Cout; ++o) {
          const T* dyo = dy + ((size_t)n * Cout + o) * Ho * Wo;
          const T* wk = Wt + ((size_t)o * Cin + c) * k * k;
          for (int kh = 0; kh < k; ++kh)
            for (int kw = 0; kw < k; ++kw) {
              const T w = wk[kh * k + kw];
              for (int ho = 0; ho < Ho; ++ho) {
                const int ih = ho * stride + kh - pad;
                if (ih < 0 || ih >= H) continue;
                for (int wo = 0; wo < Wo; ++wo) {
                  const int iw = wo * stride + kw - pad;
                  if (iw >= 0 && iw < W) dxc[(size_t)ih * W + iw] += w * dyo[ho * Wo + wo];
                }
              }
            }
        }
      }
  }
  if (dW) {
#pragma omp parallel for collapse(2) schedule(static)
    for (int o = 0; o < Cout; ++o)
      for (int c = 0; c < Cin; ++c) {
        T* wk = dW + ((size_t)o * Cin + c) * k * k;
        for (int n = 0; n < B; ++n) {
          const T* dyo = dy + ((size_t)n * Cout + o) * Ho * Wo;
          const T* xc = x + ((size_t)n * Cin + c) * H * W;
          for (int kh = 0; kh < k; ++kh)
            for (int kw = 0; kw < k; ++kw) {
              T s = 0;
              for (int ho = 0; ho < Ho; ++ho) {
                const int ih = ho * stride + kh - pad;
                if (ih < 0 || ih >= H) continue;
                for (int wo = 0; wo < Wo; ++wo) {
                  const int iw = wo * stride + kw - pad;
                  if (iw >= 0 && iw < W) s += dyo[ho * Wo + wo] * xc[(size_t)ih * W + iw];
                }
              }
              wk[kh * k + kw] += s;
            }
        }
      }
  }
  if (db)
    for (int o = 0; o < Cout; ++o) {
      T s = 0;
      for (int n = 0; n < B; ++n)
        for (int i = 0; i < Ho * Wo; ++i) s += dy[((size_t)n * Cout + o) * Ho * Wo + i];
      db[o] += s;
    }
}

struct G16Layout {  // same order as GLayout, Linear(100, 2048)
  size_t L1W, L1b, a1, C1W, C1b, g1, be1, a2, C2W, C2b, g2, be2, a3, C3W, C3b, total;
  explicit G16Layout(int C) {
    size_t o = 0;
    L1W = o; o += 2048 * 100;
    L1b = o; o += 2048;
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
struct D16Layout {  // conv branch, then dense branch, then the joint Linear (ConcatTable order, models.lua:306-313)
  size_t cW[4], cb[4], ca[4], F1W, F1b, af, E1W, E1b, ae1, E2W, E2b, ae2, JW, Jb, total;
  int cin[4], cout[4];
  explicit D16Layout(int C) {
    const int ci[4] = {C, 128, 128, 512}, co[4] = {128, 128, 512, 1024};
    size_t o = 0;
    for (int i = 0; i < 4; ++i) {
      cin[i] = ci[i]; cout[i] = co[i];
      cW[i] = o; o += (size_t)co[i] * ci[i] * 9;
      cb[i] = o; o += co[i];
      ca[i] = o; o += 1;
    }
    F1W = o; o += (size_t)1024 * 4096;
    F1b = o; o += 1024;
    af = o; o += 1;
    E1W = o; o += (size_t)128 * C * 256;
    E1b = o; o += 128;
    ae1 = o; o += 1;
    E2W = o; o += 128 * 128;
    E2b = o; o += 128;
    ae2 = o; o += 1;
    JW = o; o += 1152;
    Jb = o; o += 1;
    total = o;
  }
};
constexpr int kD16MaskPerSample = 1024 + 128;  // nn.SpatialDropout() planes (p = 0.5) + nn.Dropout() of the dense branch

template <class T>
struct G16Net {
  int B = 0, C = 3;
  std::vector<T> x, z0, h0, u0, z1, y1, h1, u1, z2, y2, h2, z3, out;
  BNSave<T> s1, s2;
  void forward(const T* P, const T* noise, int B_, int C_, T* bn_state /*768 or null*/) {
    B = B_; C = C_;
    G16Layout L(C);
    x.assign(noise, noise + (size_t)B * 100);
    z0.resize((size_t)B * 2048); h0.resize(z0.size());
    linear_fwd(B, 100, 2048, x.data(), P + L.L1W, P + L.L1b, z0.data());
    prelu_fwd(z0.size(), z0.data(), P[L.a1], h0.data());
    u0.resize((size_t)B * 128 * 64);
    up2_fwd(B, 128, 4, 4, h0.data(), u0.data());
    z1.resize((size_t)B * 256 * 64); y1.resize(z1.size()); h1.resize(z1.size());
    conv_fwd(B, 128, 8, 8, 256, 5, u0.data(), P + L.C1W, P + L.C1b, z1.data());
    bn_fwd_train(B, 256, 64, z1.data(), P + L.g1, P + L.be1, y1.data(), s1, bn_state, bn_state ? bn_state + 256 : nullptr);
    prelu_fwd(y1.size(), y1.data(), P[L.a2], h1.data());
    u1.resize((size_t)B * 256 * 256);
    up2_fwd(B, 256, 8, 8, h1.data(), u1.data());
    z2.resize((size_t)B * 128 * 256); y2.resize(z2.size()); h2.resize(z2.size());
    conv_fwd(B, 256, 16, 16, 128, 5, u1.data(), P + L.C2W, P + L.C2b, z2.data());
    bn_fwd_train(B, 128, 256, z2.data(), P + L.g2, P + L.be2, y2.data(), s2, bn_state ? bn_state + 512 : nullptr,
                 bn_state ? bn_state + 640 : nullptr);
    prelu_fwd(y2.size(), y2.data(), P[L.a3], h2.data());
    z3.resize((size_t)B * C * 256); out.resize(z3.size());
    conv_fwd(B, 128, 16, 16, C, 3, h2.data(), P + L.C3W, P + L.C3b, z3.data());
    for (size_t i = 0; i < z3.size(); ++i) out[i] = sigmoid(z3[i]);
  }
  void backward(const T* P, const T* dout, T* dP) {
    G16Layout L(C);
    std::vector<T> dz3(z3.size());
    for (size_t i = 0; i < dz3.size(); ++i) dz3[i] = dout[i] * out[i] * (T(1) - out[i]);
    std::vector<T> dh2(h2.size());
    conv_bwd(B, 128, 16, 16, C, 3, h2.data(), P + L.C3W, dz3.data(), dh2.data(), dP + L.C3W, dP + L.C3b);
    std::vector<T> dy2(y2.size()), dz2(z2.size());
    prelu_bwd(y2.size(), y2.data(), P[L.a3], dh2.data(), dy2.data(), dP + L.a3);
    bn_bwd(B, 128, 256, z2.data(), P + L.g2, s2, dy2.data(), dz2.data(), dP + L.g2, dP + L.be2);
    std::vector<T> du1(u1.size()), dh1(h1.size());
    conv_bwd(B, 256, 16, 16, 128, 5, u1.data(), P + L.C2W, dz2.data(), du1.data(), dP + L.C2W, dP + L.C2b);
    up2_bwd(B, 256, 8, 8, du1.data(), dh1.data());
    std::vector<T> dy1(y1.size()), dz1(z1.size());
    prelu_bwd(y1.size(), y1.data(), P[L.a2], dh1.data(), dy1.data(), dP + L.a2);
    bn_bwd(B, 256, 64, z1.data(), P + L.g1, s1, dy1.data(), dz1.data(), dP + L.g1, dP + L.be1);
    std::vector<T> du0(u0.size()), dh0(h0.size());
    conv_bwd(B, 128, 8, 8, 256, 5, u0.data(), P + L.C1W, dz1.data(), du0.data(), dP + L.C1W, dP + L.C1b);
    up2_bwd(B, 128, 4, 4, du0.data(), dh0.data());
    std::vector<T> dz0(z0.size());
    prelu_bwd(z0.size(), z0.data(), P[L.a1], dh0.data(), dz0.data(), dP + L.a1);
    linear_bwd(B, 100, 2048, x.data(), P + L.L1W, dz0.data(), (T*)nullptr, dP + L.L1W, dP + L.L1b);
  }
};

template <class T>
struct D16Net {
  int B = 0, C = 3;
  bool training = true;
  std::vector<T> x, z[4], h[4], p1, d3, zf, hf, ze1, he1, de1, ze2, he2, joint, logit, out, mask;
  void forward(const T* P, const T* img, int B_, int C_, bool training_, const T* masks) {
    B = B_; C = C_; training = training_;
    D16Layout L(C);
    x.assign(img, img + (size_t)B * C * 256);
    if (training) mask.assign(masks, masks + (size_t)B * kD16MaskPerSample);
    // ---- conv branch ----
    z[0].resize((size_t)B * 128 * 256); h[0].resize(z[0].size());
    conv_fwd(B, C, 16, 16, 128, 3, x.data(), P + L.cW[0], P + L.cb[0], z[0].data());
    prelu_fwd(z[0].size(), z[0].data(), P[L.ca[0]], h[0].data());
    z[1].resize(z[0].size()); h[1].resize(z[0].size());
    conv_fwd(B, 128, 16, 16, 128, 3, h[0].data(), P + L.cW[1], P + L.cb[1], z[1].data());
    prelu_fwd(z[1].size(), z[1].data(), P[L.ca[1]], h[1].data());
    p1.resize((size_t)B * 128 * 64);
    avgpool2_fwd(B * 128, 16, 16, h[1].data(), p1.data());
    z[2].resize((size_t)B * 512 * 16); h[2].resize(z[2].size());
    convs_fwd(B, 128, 8, 8, 512, 3, 2, 1, p1.data(), P + L.cW[2], P + L.cb[2], z[2].data());  // 8x8 -> 4x4
    prelu_fwd(z[2].size(), z[2].data(), P[L.ca[2]], h[2].data());
    z[3].resize((size_t)B * 1024 * 4); h[3].resize(z[3].size());
    convs_fwd(B, 512, 4, 4, 1024, 3, 2, 1, h[2].data(), P + L.cW[3], P + L.cb[3], z[3].data());  // 4x4 -> 2x2
    prelu_fwd(z[3].size(), z[3].data(), P[L.ca[3]], h[3].data());
    d3.resize(h[3].size());  // nn.SpatialDropout(): p = 0.5, one flag per plane, no rescale; evaluate(): * (1-p)
    for (int b = 0; b < B; ++b)
      for (int c = 0; c < 1024; ++c) {
        const T m = training ? mask[(size_t)b * kD16MaskPerSample + c] : T(0.5);
        for (int q = 0; q < 4; ++q) d3[((size_t)b * 1024 + c) * 4 + q] = h[3][((size_t)b * 1024 + c) * 4 + q] * m;
      }
    zf.resize((size_t)B * 1024); hf.resize(zf.size());
    linear_fwd(B, 4096, 1024, d3.data(), P + L.F1W, P + L.F1b, zf.data());
    prelu_fwd(zf.size(), zf.data(), P[L.af], hf.data());
    // ---- dense branch ----
    ze1.resize((size_t)B * 128); he1.resize(ze1.size()); de1.resize(ze1.size()); ze2.resize(ze1.size()); he2.resize(ze1.size());
    linear_fwd(B, C * 256, 128, x.data(), P + L.E1W, P + L.E1b, ze1.data());
    prelu_fwd(ze1.size(), ze1.data(), P[L.ae1], he1.data());
    for (int b = 0; b < B; ++b)
      for (int j = 0; j < 128; ++j)
        de1[(size_t)b * 128 + j] =
            training ? he1[(size_t)b * 128 + j] * mask[(size_t)b * kD16MaskPerSample + 1024 + j] * T(2) : he1[(size_t)b * 128 + j];
    linear_fwd(B, 128, 128, de1.data(), P + L.E2W, P + L.E2b, ze2.data());
    prelu_fwd(ze2.size(), ze2.data(), P[L.ae2], he2.data());
    // ---- JoinTable(2) -> Linear(1152, 1) -> Sigmoid ----
    joint.resize((size_t)B * 1152);
    for (int b = 0; b < B; ++b) {
      std::copy(hf.begin() + (size_t)b * 1024, hf.begin() + (size_t)(b + 1) * 1024, joint.begin() + (size_t)b * 1152);
      std::copy(he2.begin() + (size_t)b * 128, he2.begin() + (size_t)(b + 1) * 128, joint.begin() + (size_t)b * 1152 + 1024);
    }
    logit.resize(B); out.resize(B);
    linear_fwd(B, 1152, 1, joint.data(), P + L.JW, P + L.Jb, logit.data());
    for (int b = 0; b < B; ++b) out[b] = d_output(sigmoid(logit[b]));  // fp32 at the criterion boundary (fg_oracle.cpp)
  }
  void backward(const T* P, const T* dout, T* dP, T* dimg) {
    D16Layout L(C);
    std::vector<T> dlogit(B), djoint((size_t)B * 1152);
    for (int b = 0; b < B; ++b) dlogit[b] = dout[b] * out[b] * (T(1) - out[b]);
    linear_bwd(B, 1152, 1, joint.data(), P + L.JW, dlogit.data(), djoint.data(), dP + L.JW, dP + L.Jb);
    std::vector<T> dhf((size_t)B * 1024), dhe2((size_t)B * 128);
    for (int b = 0; b < B; ++b) {
      std::copy(djoint.begin() + (size_t)b * 1152, djoint.begin() + (size_t)b * 1152 + 1024, dhf.begin() + (size_t)b * 1024);
      std::copy(djoint.begin() + (size_t)b * 1152 + 1024, djoint.begin() + (size_t)(b + 1) * 1152, dhe2.begin() + (size_t)b * 128);
    }
    std::vector<T> dx_dense((size_t)B * C * 256), dx_conv(dx_dense.size());
    {  // dense branch
      std::vector<T> dze2(dhe2.size()), dde1(dhe2.size()), dhe1(dhe2.size()), dze1(dhe2.size());
      prelu_bwd(dze2.size(), ze2.data(), P[L.ae2], dhe2.data(), dze2.data(), dP + L.ae2);
      linear_bwd(B, 128, 128, de1.data(), P + L.E2W, dze2.data(), dde1.data(), dP + L.E2W, dP + L.E2b);
      for (int b = 0; b < B; ++b)
        for (int j = 0; j < 128; ++j)
          dhe1[(size_t)b * 128 + j] =
              training ? dde1[(size_t)b * 128 + j] * mask[(size_t)b * kD16MaskPerSample + 1024 + j] * T(2) : dde1[(size_t)b * 128 + j];
      prelu_bwd(dze1.size(), ze1.data(), P[L.ae1], dhe1.data(), dze1.data(), dP + L.ae1);
      linear_bwd(B, C * 256, 128, x.data(), P + L.E1W, dze1.data(), dx_dense.data(), dP + L.E1W, dP + L.E1b);
    }
    {  // conv branch
      std::vector<T> dzf(dhf.size()), dd3((size_t)B * 4096), dh3(dd3.size()), dz3(dd3.size());
      prelu_bwd(dzf.size(), zf.data(), P[L.af], dhf.data(), dzf.data(), dP + L.af);
      linear_bwd(B, 4096, 1024, d3.data(), P + L.F1W, dzf.data(), dd3.data(), dP + L.F1W, dP + L.F1b);
      for (int b = 0; b < B; ++b)
        for (int c = 0; c < 1024; ++c) {
          const T m = training ? mask[(size_t)b * kD16MaskPerSample + c] : T(0.5);
          for (int q = 0; q < 4; ++q) dh3[((size_t)b * 1024 + c) * 4 + q] = dd3[((size_t)b * 1024 + c) * 4 + q] * m;
        }
      prelu_bwd(dz3.size(), z[3].data(), P[L.ca[3]], dh3.data(), dz3.data(), dP + L.ca[3]);
      std::vector<T> dh2(h[2].size()), dz2(h[2].size());
      convs_bwd(B, 512, 4, 4, 1024, 3, 2, 1, h[2].data(), P + L.cW[3], dz3.data(), dh2.data(), dP + L.cW[3], dP + L.cb[3]);
      prelu_bwd(dz2.size(), z[2].data(), P[L.ca[2]], dh2.data(), dz2.data(), dP + L.ca[2]);
      std::vector<T> dp1(p1.size()), dh1(h[1].size()), dz1(h[1].size());
      convs_bwd(B, 128, 8, 8, 512, 3, 2, 1, p1.data(), P + L.cW[2], dz2.data(), dp1.data(), dP + L.cW[2], dP + L.cb[2]);
      avgpool2_bwd(B * 128, 16, 16, dp1.data(), dh1.data());
      prelu_bwd(dz1.size(), z[1].data(), P[L.ca[1]], dh1.data(), dz1.data(), dP + L.ca[1]);
      std::vector<T> dh0(h[0].size()), dz0(h[0].size());
      conv_bwd(B, 128, 16, 16, 128, 3, h[0].data(), P + L.cW[1], dz1.data(), dh0.data(), dP + L.cW[1], dP + L.cb[1]);
      prelu_bwd(dz0.size(), z[0].data(), P[L.ca[0]], dh0.data(), dz0.data(), dP + L.ca[0]);
      conv_bwd(B, C, 16, 16, 128, 3, x.data(), P + L.cW[0], dz0.data(), dx_conv.data(), dP + L.cW[0], dP + L.cb[0]);
    }
    if (dimg)  // nn.ConcatTable backward: the input gradient is the sum over the branches
      for (size_t i = 0; i < dx_conv.size(); ++i) dimg[i] = dx_conv[i] + dx_dense[i];
  }
};
