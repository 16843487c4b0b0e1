// TEST INFRASTRUCTURE (see fg_oracle.cpp header): CPU restatement of the coarse-to-fine nets and loop.
// PARITY UNPINNED: the reference has no tests / golden vectors and Torch7 cannot run here.
// Included by fg_oracle.cpp inside its anonymous namespace (uses conv_fwd/conv_bwd/linear_*/prelu_*/bce_*/adam).
//
//   G = models_c2f.lua:113-145 create_G_d   (JoinTable{noise, coarse} -> 5 SpatialConvolutionUpsample(factor 1))
//   D = models_c2f.lua:237-278 create_D_c   (CAddTable{diff, coarse} -> 4 conv/PReLU, 2 MaxPool, Dropout, 2 Linear)
//   loop = adversarial_c2f.lua:121-187      (fevalD :40-81, fevalG_on_D :85-116, stock optim.adam)
//
// layers/cudnnSpatialConvolutionUpsample.lua:4-16 with factor = 1 is a plain "same" convolution (pad (k-1)/2)
// followed by a view that changes nothing (:23), so conv_fwd/conv_bwd restate it.
#pragma once

// nn.SpatialMaxPooling(2,2): kW=kH=2, stride defaults to the kernel size.  THNN scans the window row-major and
// keeps the first strict maximum (`val > maxval`), the backward routes the gradient to that element.
template <class T>
void maxpool2_fwd(int BC, int H, int W, const T* x, T* y, unsigned char* arg) {
  const int Ho = H / 2, Wo = W / 2;
#pragma omp parallel for schedule(static)
  for (int nc = 0; nc < BC; ++nc)
    for (int h = 0; h < Ho; ++h)
      for (int w = 0; w < Wo; ++w) {
        const T* r0 = x + ((size_t)nc * H + 2 * h) * W + 2 * w;
        const T v[4] = {r0[0], r0[1], r0[W], r0[W + 1]};
        int best = 0;
        for (int j = 1; j < 4; ++j)
          if (v[j] > v[best]) best = j;
        const size_t o = ((size_t)nc * Ho + h) * Wo + w;
        y[o] = v[best];
        if (arg) arg[o] = (unsigned char)best;
      }
}
template <class T>
void maxpool2_bwd(int BC, int H, int W, const T* dy, const unsigned char* arg, T* dx) {
  const int Ho = H / 2, Wo = W / 2;
  std::fill(dx, dx + (size_t)BC * H * W, T(0));
#pragma omp parallel for schedule(static)
  for (int nc = 0; nc < BC; ++nc)
    for (int h = 0; h < Ho; ++h)
      for (int w = 0; w < Wo; ++w) {
        const size_t o = ((size_t)nc * Ho + h) * Wo + w;
        const int j = arg[o];
        dx[((size_t)nc * H + 2 * h + (j >> 1)) * W + 2 * w + (j & 1)] = dy[o];
      }
}

// getParameters() order: module order, weight then bias; nn.PReLU() = one shared slope
struct C2fGLayout {  // models_c2f.lua:124-133
  int64_t cW[5], cb[5], ca[4], total;
  int cin[5], cout[5], k[5];
  explicit C2fGLayout(int C) {
    const int ci[5] = {C + 1, 64, 64, 128, 256}, co[5] = {64, 64, 128, 256, C}, kk[5] = {3, 3, 5, 5, 7};
    int64_t o = 0;
    for (int i = 0; i < 5; ++i) {
      cin[i] = ci[i]; cout[i] = co[i]; k[i] = kk[i];
      cW[i] = o; o += (int64_t)co[i] * ci[i] * kk[i] * kk[i];
      cb[i] = o; o += co[i];
      if (i < 4) { ca[i] = o; o += 1; }
    }
    total = o;
  }
};
struct C2fDLayout {  // models_c2f.lua:247-265
  int64_t cW[4], cb[4], ca[4], L1W, L1b, a5, L2W, L2b, total;
  int cin[4], cout[4];
  explicit C2fDLayout(int C) {
    const int ci[4] = {C, 64, 64, 128}, co[4] = {64, 64, 128, 256};
    int64_t o = 0;
    for (int i = 0; i < 4; ++i) {
      cin[i] = ci[i]; cout[i] = co[i];
      cW[i] = o; o += (int64_t)co[i] * ci[i] * 9;
      cb[i] = o; o += co[i];
      ca[i] = o; o += 1;
    }
    L1W = o; o += (int64_t)512 * 16384;
    L1b = o; o += 512;
    a5 = o; o += 1;
    L2W = o; o += 512;
    L2b = o; o += 1;
    total = o;
  }
};
constexpr int kC2fMaskPerSample = 16384 + 512;  // nn.Dropout keep flags: [256][8][8] (:258) then [512] (:264)

template <class T>
struct C2fGNet {
  int B = 0, C = 3;
  std::vector<T> x;              // JoinTable(2,2): [B][1+C][32][32], noise plane first (models_c2f.lua:116)
  std::vector<T> z[5], h[4];     // conv outputs, PReLU outputs; z[4] is the generated diff
  void forward(const T* P, const T* noise, const T* cond, int B_, int C_) {
    B = B_; C = C_;
    C2fGLayout L(C);
    x.resize((size_t)B * (C + 1) * 1024);
    for (int b = 0; b < B; ++b) {
      std::copy(noise + (size_t)b * 1024, noise + (size_t)(b + 1) * 1024, x.begin() + (size_t)b * (C + 1) * 1024);
      std::copy(cond + (size_t)b * C * 1024, cond + (size_t)(b + 1) * C * 1024,
                x.begin() + (size_t)b * (C + 1) * 1024 + 1024);
    }
    const T* cur = x.data();
    for (int i = 0; i < 5; ++i) {
      const size_t n = (size_t)B * L.cout[i] * 1024;
      z[i].resize(n);
      conv_fwd(B, L.cin[i], 32, 32, L.cout[i], L.k[i], cur, P + L.cW[i], P + L.cb[i], z[i].data());
      if (i < 4) {
        h[i].resize(n);
        prelu_fwd(n, z[i].data(), P[L.ca[i]], h[i].data());
        cur = h[i].data();
      }
    }
  }
  // dout [B][C][32][32]; accumulates into dP; the gradient w.r.t. {noise, coarse} is not needed by the loop
  void backward(const T* P, const T* dout, T* dP) {
    C2fGLayout L(C);
    std::vector<T> dz(dout, dout + (size_t)B * C * 1024), dh;
    for (int i = 4; i >= 0; --i) {
      const T* in = i == 0 ? x.data() : h[i - 1].data();
      const bool need_dx = i > 0;
      if (need_dx) dh.assign((size_t)B * L.cin[i] * 1024, T(0));
      conv_bwd(B, L.cin[i], 32, 32, L.cout[i], L.k[i], in, P + L.cW[i], dz.data(), need_dx ? dh.data() : nullptr,
               dP + L.cW[i], dP + L.cb[i]);
      if (i > 0) {
        dz.resize(dh.size());
        prelu_bwd(dh.size(), z[i - 1].data(), P[L.ca[i - 1]], dh.data(), dz.data(), dP + L.ca[i - 1]);
      }
    }
  }
};

template <class T>
struct C2fDNet {
  int B = 0, C = 3;
  bool training = true;
  std::vector<T> x;                    // CAddTable: diff + coarse (models_c2f.lua:240)
  std::vector<T> z[4], h[4], p2, p4, d4;
  std::vector<unsigned char> arg2, arg4;
  std::vector<T> zl1, al1, hl1, logit, out, mask;
  void forward(const T* P, const T* diff, const T* cond, int B_, int C_, bool training_, const T* masks) {
    B = B_; C = C_; training = training_;
    C2fDLayout L(C);
    const int hw[4] = {32, 32, 16, 16};
    x.resize((size_t)B * C * 1024);
    for (size_t i = 0; i < x.size(); ++i) x[i] = diff[i] + cond[i];
    if (training) mask.assign(masks, masks + (size_t)B * kC2fMaskPerSample);
    const T* cur = x.data();
    for (int i = 0; i < 4; ++i) {
      const int H = hw[i];
      const size_t n = (size_t)B * L.cout[i] * H * H;
      z[i].resize(n); h[i].resize(n);
      conv_fwd(B, L.cin[i], H, H, L.cout[i], 3, cur, P + L.cW[i], P + L.cb[i], z[i].data());
      prelu_fwd(n, z[i].data(), P[L.ca[i]], h[i].data());
      cur = h[i].data();
      if (i == 1) {
        p2.resize(n / 4); arg2.resize(n / 4);
        maxpool2_fwd(B * 64, 32, 32, h[1].data(), p2.data(), arg2.data());
        cur = p2.data();
      } else if (i == 3) {
        p4.resize(n / 4); arg4.resize(n / 4);
        maxpool2_fwd(B * 256, 16, 16, h[3].data(), p4.data(), arg4.data());
      }
    }
    // nn.Dropout() p=0.5 (v2): train y = x*mask/(1-p), eval identity; then View(16384) in (c,h,w) order
    d4.resize(p4.size());
    for (int b = 0; b < B; ++b)
      for (int j = 0; j < 16384; ++j)
        d4[(size_t)b * 16384 + j] =
            training ? p4[(size_t)b * 16384 + j] * mask[(size_t)b * kC2fMaskPerSample + j] * T(2) : p4[(size_t)b * 16384 + j];
    zl1.resize((size_t)B * 512); al1.resize(zl1.size()); hl1.resize(zl1.size());
    linear_fwd(B, 16384, 512, d4.data(), P + L.L1W, P + L.L1b, zl1.data());
    prelu_fwd(zl1.size(), zl1.data(), P[L.a5], al1.data());
    for (int b = 0; b < B; ++b)
      for (int j = 0; j < 512; ++j)
        hl1[(size_t)b * 512 + j] = training ? al1[(size_t)b * 512 + j] * mask[(size_t)b * kC2fMaskPerSample + 16384 + j] * T(2)
                                           : al1[(size_t)b * 512 + j];
    logit.resize(B); out.resize(B);
    linear_fwd(B, 512, 1, hl1.data(), P + L.L2W, P + L.L2b, logit.data());
    for (int b = 0; b < B; ++b) out[b] = d_output(sigmoid(logit[b]));  // fp32 at the criterion boundary (fg_oracle.cpp)
  }
  // dout [B] = dLoss/d(sigmoid output); dP may be null (weight grads skipped); ddiff = MODEL_D.gradInput[1]
  void backward(const T* P, const T* dout, T* dP, T* ddiff) {
    C2fDLayout L(C);
    const int hw[4] = {32, 32, 16, 16};
    T dummy = 0;
    auto gw = [&](int64_t off) { return dP ? dP + off : (T*)nullptr; };   // weight/bias grads (skipped if null)
    auto ga = [&](int64_t off) { return dP ? dP + off : &dummy; };        // PReLU slope grads
    std::vector<T> dlogit(B);
    for (int b = 0; b < B; ++b) dlogit[b] = dout[b] * out[b] * (T(1) - out[b]);
    std::vector<T> dhl1((size_t)B * 512), dal1(dhl1.size()), dzl1(dhl1.size());
    linear_bwd(B, 512, 1, hl1.data(), P + L.L2W, dlogit.data(), dhl1.data(), gw(L.L2W), gw(L.L2b));
    for (int b = 0; b < B; ++b)
      for (int j = 0; j < 512; ++j)
        dal1[(size_t)b * 512 + j] = training ? dhl1[(size_t)b * 512 + j] * mask[(size_t)b * kC2fMaskPerSample + 16384 + j] * T(2)
                                            : dhl1[(size_t)b * 512 + j];
    prelu_bwd(dzl1.size(), zl1.data(), P[L.a5], dal1.data(), dzl1.data(), ga(L.a5));
    std::vector<T> dd4((size_t)B * 16384), dp4(dd4.size());
    linear_bwd(B, 16384, 512, d4.data(), P + L.L1W, dzl1.data(), dd4.data(), gw(L.L1W), gw(L.L1b));
    for (int b = 0; b < B; ++b)
      for (int j = 0; j < 16384; ++j)
        dp4[(size_t)b * 16384 + j] =
            training ? dd4[(size_t)b * 16384 + j] * mask[(size_t)b * kC2fMaskPerSample + j] * T(2) : dd4[(size_t)b * 16384 + j];
    std::vector<T> dh, dz, dx;
    for (int i = 3; i >= 0; --i) {
      const int H = hw[i];
      const size_t n = (size_t)B * L.cout[i] * H * H;
      if (i == 3) {
        dh.resize(n);
        maxpool2_bwd(B * 256, 16, 16, dp4.data(), arg4.data(), dh.data());
      } else if (i == 1) {
        dh.resize(n);
        maxpool2_bwd(B * 64, 32, 32, dx.data(), arg2.data(), dh.data());
      } else {
        dh.swap(dx);
      }
      dz.resize(n);
      prelu_bwd(n, z[i].data(), P[L.ca[i]], dh.data(), dz.data(), ga(L.ca[i]));
      const T* in = i == 0 ? x.data() : (i == 2 ? p2.data() : h[i - 1].data());
      const bool need_dx = i > 0 || ddiff != nullptr;
      if (need_dx) dx.assign((size_t)B * L.cin[i] * H * H, T(0));
      conv_bwd(B, L.cin[i], H, H, L.cout[i], 3, in, P + L.cW[i], dz.data(), need_dx ? dx.data() : nullptr, gw(L.cW[i]),
               gw(L.cb[i]));
    }
    if (ddiff) std::copy(dx.begin(), dx.end(), ddiff);  // CAddTable backward: identity to both addends
  }
};

// One iteration of the adversarial_c2f.lua loop body (D_iterations = G_iterations = 1, adam for both).
//   real_diff[B/2,C,32,32]  fine-minus-coarse of the real half            (:127-130)
//   condD[B,C,32,32]        coarse images: rows < B/2 belong to the real samples, the rest to the fakes (:129,:137-141)
//   noiseD[B/2,1,32,32]     U(-1,1) for the generated half                (:135, :145)
//   condG[B,C,32,32], noiseG[B,1,32,32]  redrawn for the G step           (:168-174)
//   masksD/masksG[B,16896]  nn.Dropout keep flags of the two D forwards
// `optim.adam` (un-pinned third-party, 2015) is the routine interruptable_optimizers.lua:49-94 was copied from;
// the restatement uses the same update.
template <class T>
void c2f_train_iteration(int B, int C, const Hyper& hp, const T* real_diff, const T* condD, const T* noiseD,
                         const T* condG, const T* noiseG, const T* masksD, const T* masksG, T* PD, T* PG, T* mD, T* vD,
                         T* mG, T* vG, int* tD, int* tG, double* stats, T* gradD_out, T* gradG_out, T* fake_out,
                         T* outD_out) {
  C2fGLayout LG(C);
  C2fDLayout LD(C);
  const int Bh = B / 2;
  const size_t img = (size_t)C * 1024;
  C2fGNet<T> G;
  C2fDNet<T> D;
  // ---- D step ----
  G.forward(PG, noiseD, condD + Bh * img, Bh, C);
  if (fake_out) std::copy(G.z[4].begin(), G.z[4].end(), fake_out);
  std::vector<T> inputs((size_t)B * img), targets(B);
  std::copy(real_diff, real_diff + Bh * img, inputs.begin());
  std::copy(G.z[4].begin(), G.z[4].end(), inputs.begin() + Bh * img);
  for (int i = 0; i < B; ++i) targets[i] = i < Bh ? T(1) : T(0);
  std::vector<T> gD(LD.total, T(0));
  D.forward(PD, inputs.data(), condD, B, C, true, masksD);
  if (outD_out) std::copy(D.out.begin(), D.out.end(), outD_out);
  T fD = bce_fwd(B, D.out.data(), targets.data());
  std::vector<T> df(B);
  bce_bwd(B, D.out.data(), targets.data(), df.data());
  D.backward(PD, df.data(), gD.data(), nullptr);
  fD += penalty_clamp(LD.total, PD, gD.data(), T(hp.D_L1), T(hp.D_L1), T(hp.D_L2), T(hp.D_clamp));  // :56-63, :74-76
  double conf[4] = {0, 0, 0, 0};
  for (int i = 0; i < B; ++i) {
    const bool pred1 = D.out[i] > T(0.5);
    const bool t1 = i < Bh;
    conf[(pred1 ? 0 : 1) + (t1 ? 0 : 2)] += 1;
  }
  if (gradD_out) std::copy(gD.begin(), gD.end(), gradD_out);
  *tD += 1;
  adam(LD.total, PD, gD.data(), mD, vD, *tD, hp.lr_D, hp.beta1, hp.beta2, hp.eps);
  // ---- G step ----
  std::vector<T> gG(LG.total, T(0));
  G.forward(PG, noiseG, condG, B, C);
  for (int i = 0; i < B; ++i) targets[i] = T(1);
  D.forward(PD, G.z[4].data(), condG, B, C, true, masksG);
  T fG = bce_fwd(B, D.out.data(), targets.data());
  bce_bwd(B, D.out.data(), targets.data(), df.data());
  std::vector<T> ddiff((size_t)B * img);
  D.backward(PD, df.data(), nullptr, ddiff.data());
  G.backward(PG, ddiff.data(), gG.data());
  // same quirk as adversarial.lua:223: sign(p) is scaled by G_L2 (adversarial_c2f.lua:108)
  fG += penalty_clamp(LG.total, PG, gG.data(), T(hp.G_L1), T(hp.G_L2), T(hp.G_L2), T(hp.G_clamp));
  if (gradG_out) std::copy(gG.begin(), gG.end(), gradG_out);
  *tG += 1;
  adam(LG.total, PG, gG.data(), mG, vG, *tG, hp.lr_G, hp.beta1, hp.beta2, hp.eps);
  stats[0] = (double)fD; stats[1] = (double)fG;
  stats[2] = conf[0]; stats[3] = conf[1]; stats[4] = conf[2]; stats[5] = conf[3];
  stats[6] = 0; stats[7] = 0;
}
