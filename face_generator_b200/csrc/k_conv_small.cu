// Convolutions with a tiny channel count on one side (the 3-channel image end of both networks):
//   G.C3  128 -> C(1|3) 3x3   (models.lua:73)      forward = small-N, dgrad = small-K, wgrad = small/big
//   D.C1  C(1|3) -> 64  3x3   (models.lua:385)     forward = small-K, dgrad = small-N, wgrad = small/big
// These are NOT dense contractions (K = 27 or N = 3): they are bound by the HBM traffic of the big
// activation tensor, so they get bandwidth-shaped kernels instead of GEMM tiles:
// coalesced channel-fastest accesses, the small operand broadcast through L1/smem, warp-shuffle reductions.
// All tensors NHWC fp32; stride 1, pad (k-1)/2, k*k*Cs <= 36.
#include "fg_internal.h"

#define LAUNCH_CHECK(c)                 \
  do {                                  \
    (c)->launches++;                    \
    FG_CUDA(cudaGetLastError());        \
  } while (0)

namespace {
constexpr int kMaxSmallW = 36 * 128;  // k*k*Cs * Cb floats of weights in shared memory

// ---- small contraction: out[p][n] = bias[n] + sum_{t,c<Cs} in[pix(p,t)][c] * Wp[t][n][c] ---------------
// one thread per (pixel, n), n fastest: the Cs inputs of a pixel are a warp broadcast, writes are coalesced
template <int CS>
__global__ void __launch_bounds__(256) conv_smallk_kernel(const float* __restrict__ in, const float* __restrict__ Wp,
                                                          const float* __restrict__ bias, float* __restrict__ out, int B,
                                                          int H, int W, int N, int k) {
  __shared__ float ws[kMaxSmallW];  // [t][c][n]
  const int KK = k * k, pad = (k - 1) / 2;
  for (int i = threadIdx.x; i < KK * CS * N; i += blockDim.x) {
    const int n = i % N, c = (i / N) % CS, t = i / (N * CS);
    ws[i] = Wp[((int64_t)t * N + n) * CS + c];
  }
  __syncthreads();
  const int64_t total = (int64_t)B * H * W * N;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(i % N);
    int64_t p = i / N;
    const int x = (int)(p % W);
    p /= W;
    const int y = (int)(p % H);
    const int b = (int)(p / H);
    float acc = bias ? bias[n] : 0.f;
    for (int t = 0; t < KK; ++t) {
      const int iy = y + t / k - pad, ix = x + t % k - pad;
      if (iy < 0 || iy >= H || ix < 0 || ix >= W) continue;
      const float* ip = in + (((int64_t)b * H + iy) * W + ix) * CS;
#pragma unroll
      for (int c = 0; c < CS; ++c) acc = fmaf(__ldg(ip + c), ws[(t * CS + c) * N + n], acc);
    }
    out[i] = acc;
  }
}

// ---- small output: out[p][n<NS] = bias[n] + sum_{t,c} in[pix(p,t)][c] * Wp[t][n][c] ---------------------
// one warp per pixel, lane owns VEC consecutive channels (C = 32*VEC), warp-shuffle reduction of NS sums
template <int NS, int VEC>
__global__ void __launch_bounds__(256) conv_smalln_kernel(const float* __restrict__ in, const float* __restrict__ Wp,
                                                          const float* __restrict__ bias, float* __restrict__ out, int B,
                                                          int H, int W, int k) {
  constexpr int C = 32 * VEC;
  __shared__ float ws[kMaxSmallW];  // [t][n][c]
  const int KK = k * k, pad = (k - 1) / 2;
  for (int i = threadIdx.x; i < KK * NS * C; i += blockDim.x) ws[i] = Wp[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t P = (int64_t)B * H * W;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t p = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; p < P; p += warps) {
    const int x = (int)(p % W);
    const int y = (int)((p / W) % H);
    const int b = (int)(p / ((int64_t)W * H));
    float acc[NS];
#pragma unroll
    for (int n = 0; n < NS; ++n) acc[n] = 0.f;
    for (int t = 0; t < KK; ++t) {
      const int iy = y + t / k - pad, ix = x + t % k - pad;
      if (iy < 0 || iy >= H || ix < 0 || ix >= W) continue;
      const float* ip = in + (((int64_t)b * H + iy) * W + ix) * C + lane * VEC;
      float v[VEC];
      if (VEC == 4) {
        const float4 q = *reinterpret_cast<const float4*>(ip);
        v[0] = q.x; v[1] = q.y; v[2 % VEC] = q.z; v[3 % VEC] = q.w;
      } else if (VEC == 2) {
        const float2 q = *reinterpret_cast<const float2*>(ip);
        v[0] = q.x; v[1 % VEC] = q.y;
      } else {
#pragma unroll
        for (int j = 0; j < VEC; ++j) v[j] = ip[j];
      }
#pragma unroll
      for (int n = 0; n < NS; ++n) {
        const float* wp = ws + (t * NS + n) * C + lane * VEC;
#pragma unroll
        for (int j = 0; j < VEC; ++j) acc[n] = fmaf(v[j], wp[j], acc[n]);
      }
    }
#pragma unroll
    for (int n = 0; n < NS; ++n) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc[n] += __shfl_xor_sync(0xffffffffu, acc[n], o);
    }
    if (lane == 0) {
#pragma unroll
      for (int n = 0; n < NS; ++n) out[p * NS + n] = acc[n] + (bias ? bias[n] : 0.f);
    }
  }
}

// ---- weight gradient with one small and one big side ----------------------------------------------------
//   G.C3:  dW[n<Cs][c][t] = sum_p dY[p][n] * X[p+off_t][c]      big = X  (Cb = 128), small = dY, sign = -1
//   D.C1:  dW[n][c<Cs][t] = sum_p dY[p][n] * X[p+off_t][c]      big = dY (Cb = 64),  small = X,  sign = +1
// One thread per big channel walks along image rows with a 3x3 sliding window of the small tensor held in
// registers (staged per row through shared memory as float4): 3 broadcast LDS.128 + 1 coalesced LDG per
// 9*Cs FMAs.  Blocks write per-block partial sums; a second tiny kernel reduces them in a fixed order
// (deterministic, no atomics).
//   window index idx = r*3 + c holds small[y+r-1][x+c-1];  tap t = idx (sign +1) or 8 - idx (sign -1)
template <int CS>
__global__ void __launch_bounds__(128) wgrad_smallbig_kernel(const float* __restrict__ big, const float* __restrict__ small,
                                                             float* __restrict__ part, int B, int H, int W, int Cb,
                                                             int rows_per_block) {
  __shared__ float4 sm[3][68];
  const int lanes = 128 / Cb;
  const int cb = threadIdx.x % Cb, pl = threadIdx.x / Cb;
  const int xs = W / lanes, x_begin = pl * xs, x_end = x_begin + xs;
  float acc[9][CS];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int c = 0; c < CS; ++c) acc[t][c] = 0.f;
  const int BH = B * H;
  const int r0 = blockIdx.x * rows_per_block, r1 = min(BH, r0 + rows_per_block);
  for (int r = r0; r < r1; ++r) {
    const int b = r / H, y = r - b * H;
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * (W + 2); i += 128) {
      const int rr = i / (W + 2), cc = i - rr * (W + 2);
      const int yy = y + rr - 1, xx = cc - 1;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
        const float* sp = small + (((int64_t)b * H + yy) * W + xx) * CS;
        v.x = sp[0];
        if (CS > 1) v.y = sp[1 % CS];
        if (CS > 2) v.z = sp[2 % CS];
        if (CS > 3) v.w = sp[3 % CS];
      }
      sm[rr][cc] = v;
    }
    __syncthreads();
    const float* bp = big + (int64_t)r * W * Cb + cb;
    float4 w0[3], w1[3], w2[3];
#pragma unroll
    for (int rr = 0; rr < 3; ++rr) {
      w0[rr] = sm[rr][x_begin];
      w1[rr] = sm[rr][x_begin + 1];
    }
    for (int x = x_begin; x < x_end; ++x) {
      const float bv = bp[(int64_t)x * Cb];
#pragma unroll
      for (int rr = 0; rr < 3; ++rr) {
        w2[rr] = sm[rr][x + 2];
        const float4 q[3] = {w0[rr], w1[rr], w2[rr]};
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {
          acc[rr * 3 + cc][0] = fmaf(bv, q[cc].x, acc[rr * 3 + cc][0]);
          if (CS > 1) acc[rr * 3 + cc][1 % CS] = fmaf(bv, q[cc].y, acc[rr * 3 + cc][1 % CS]);
          if (CS > 2) acc[rr * 3 + cc][2 % CS] = fmaf(bv, q[cc].z, acc[rr * 3 + cc][2 % CS]);
          if (CS > 3) acc[rr * 3 + cc][3 % CS] = fmaf(bv, q[cc].w, acc[rr * 3 + cc][3 % CS]);
        }
        w0[rr] = w1[rr];
        w1[rr] = w2[rr];
      }
    }
  }
  float* dst = part + ((int64_t)blockIdx.x * lanes + pl) * (9 * CS * Cb);
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int c = 0; c < CS; ++c) dst[(t * CS + c) * Cb + cb] = acc[t][c];
}
// out (overwritten) = sum over partial rows, mapped to the packed [t][n][c] layout
__global__ void wgrad_small_reduce_kernel(const float* __restrict__ part, float* __restrict__ out, int nparts, int CS, int Cb,
                                          int sign, int transposed) {
  const int total = 9 * CS * Cb;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= total) return;
  float s = 0.f;
  for (int i = 0; i < nparts; ++i) s += part[(int64_t)i * total + j];
  const int idx = j / (CS * Cb), cs = (j / Cb) % CS, cb = j % Cb;
  const int t = sign > 0 ? idx : 8 - idx;
  out[transposed ? ((int64_t)t * Cb + cb) * CS + cs : ((int64_t)t * CS + cs) * Cb + cb] = s;
}
}  // namespace

bool k_small_eligible(const ConvGeom& g) {
  const int cs = g.Cin < g.Cout ? g.Cin : g.Cout, cb = g.Cin < g.Cout ? g.Cout : g.Cin;
  return g.ups == 1 && g.k == 3 && cs >= 1 && cs <= 4 && (cb == 32 || cb == 64 || cb == 128) && g.W <= 64 && g.W % 4 == 0;
}

// forward-type conv (also used for dgrad with the flipped/transposed pack): picks small-K or small-N
int k_conv_small(fg_ctx* c, const float* in, const float* Wp, const float* bias, float* out, ConvGeom g) {
  const int64_t P = (int64_t)g.B * g.H * g.W;
  if (g.Cin <= 4) {  // small contraction
    const int64_t total = P * g.Cout;
    int grid = (int)std::min<int64_t>((total + 255) / 256, c->sm_count * 16);
    switch (g.Cin) {
      case 1: conv_smallk_kernel<1><<<grid, 256, 0, c->stream>>>(in, Wp, bias, out, g.B, g.H, g.W, g.Cout, g.k); break;
      case 2: conv_smallk_kernel<2><<<grid, 256, 0, c->stream>>>(in, Wp, bias, out, g.B, g.H, g.W, g.Cout, g.k); break;
      case 3: conv_smallk_kernel<3><<<grid, 256, 0, c->stream>>>(in, Wp, bias, out, g.B, g.H, g.W, g.Cout, g.k); break;
      default: conv_smallk_kernel<4><<<grid, 256, 0, c->stream>>>(in, Wp, bias, out, g.B, g.H, g.W, g.Cout, g.k); break;
    }
  } else {  // small output
    int grid = (int)std::min<int64_t>((P + 7) / 8, c->sm_count * 8);
#define SN(NS_, VEC_) conv_smalln_kernel<NS_, VEC_><<<grid, 256, 0, c->stream>>>(in, Wp, bias, out, g.B, g.H, g.W, g.k)
    const int vec = g.Cin / 32;
    if (vec == 4) {
      if (g.Cout == 1) SN(1, 4); else if (g.Cout == 2) SN(2, 4); else if (g.Cout == 3) SN(3, 4); else SN(4, 4);
    } else if (vec == 2) {
      if (g.Cout == 1) SN(1, 2); else if (g.Cout == 2) SN(2, 2); else if (g.Cout == 3) SN(3, 2); else SN(4, 2);
    } else {
      if (g.Cout == 1) SN(1, 1); else if (g.Cout == 2) SN(2, 1); else if (g.Cout == 3) SN(3, 1); else SN(4, 1);
    }
#undef SN
  }
  LAUNCH_CHECK(c);
  return FG_OK;
}

// dWp[t][n][c] (overwritten) = sum_p dY[p][n] * X[pix(p,t)][c]   with min(Cin,Cout) <= 4
int k_wgrad_small(fg_ctx* c, const float* in, const float* dY, float* dWp, ConvGeom g) {
  const bool small_out = g.Cout <= 4;  // G.C3: small = dY, big = X (window read mirrored: sign -1)
  const float* big = small_out ? in : dY;
  const float* small = small_out ? dY : in;
  const int Cb = small_out ? g.Cin : g.Cout, Cs = small_out ? g.Cout : g.Cin;
  const int lanes = 128 / Cb;
  const int BH = g.B * g.H;
  int rpb = 8;
  int nblocks = (BH + rpb - 1) / rpb;
  while (nblocks * lanes > kSmallMaxParts) {
    rpb *= 2;
    nblocks = (BH + rpb - 1) / rpb;
  }
  const int sign = small_out ? -1 : 1, transposed = small_out ? 0 : 1;
#define WG(CS_) wgrad_smallbig_kernel<CS_><<<nblocks, 128, 0, c->stream>>>(big, small, c->small_ws, g.B, g.H, g.W, Cb, rpb)
  switch (Cs) {
    case 1: WG(1); break;
    case 2: WG(2); break;
    case 3: WG(3); break;
    default: WG(4); break;
  }
#undef WG
  LAUNCH_CHECK(c);
  const int total = 9 * Cs * Cb;
  wgrad_small_reduce_kernel<<<(total + 127) / 128, 128, 0, c->stream>>>(c->small_ws, dWp, nblocks * lanes, Cs, Cb, sign, transposed);
  LAUNCH_CHECK(c);
  return FG_OK;
}
