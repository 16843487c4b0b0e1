// Device kernels of the tcgen05 convolution path (included by k_conv_tc.cu after the PTX wrappers).
//
// Numerics: the tensor core accumulates into TMEM with truncation, so a long K loop into ONE accumulator
// drifts (measured 2-3e-5 relative at K = 2304, growing ~linearly with K).  Both kernels therefore
// accumulate only kChunk K-blocks (4 x 32 channels or pixels, x3 MMAs) per TMEM buffer and let the epilogue
// warps "promote" every chunk into fp32 registers with round-to-nearest adds (two TMEM buffers ping-pong, so
// the promotion overlaps the next chunk's MMAs).  The same decoupling makes the forward kernel persistent:
// while the epilogue stores tile i the MMA warp is already issuing tile i+1.
#pragma once

// K-blocks accumulated in the tensor core before promotion to registers: runtime (params.chunk, default 4;
// FG_TC_CHUNK overrides for experiments).  8 is ~3% faster but doubles the truncation drift per chunk.

struct FwdTile {
  int ph, b0, y0, x0, n0;
};
template <int BN>
__device__ __forceinline__ FwdTile fwd_decode(const TcFwdParams& p, int tile) {
  const int ntn = p.Cout / BN;
  FwdTile t;
  const int mt = tile / ntn;
  t.n0 = (tile - mt * ntn) * BN;  // n fastest: CTAs running together share the activation tile in L2
  int r = mt / p.nphase;  // the 4 output phases of one pixel tile run back to back: they share the input boxes in L2
  t.ph = mt - r * p.nphase;
  if (p.bb == 1) {
    const int per_img = p.tiles_x * p.tiles_y;
    t.b0 = r / per_img;
    r -= t.b0 * per_img;
    t.y0 = (r / p.tiles_x) * p.bh;
    t.x0 = (r % p.tiles_x) * p.bw;
  } else {
    t.b0 = r * p.bb;
    t.y0 = 0;
    t.x0 = 0;
  }
  return t;
}

// ------------------------------------------------------------------------------------------------
// tapconv: forward / dgrad.  Persistent: CTA i handles tiles i, i+gridDim.x, ...
// ------------------------------------------------------------------------------------------------
// HALO = true ("haloed tile" operand feed, 16x16 / 32x32 images, k <= 5): the M tile is 8 pixels wide x 16 rows of ONE
// image and, per 32-channel chunk, ONE TMA box of (8 + 2 pad) x (16 + 2 pad) pixels lands in shared memory (hi and lo);
// the A operand of tap (dy, dx) is the shifted 128-row window { (yi + dy + pad) * pitch + xi + dx + pad } of that box --
// the 128B-swizzle XOR is derived from the absolute shared-memory address, so a window that starts at any 128-byte
// row is a valid K-major operand with SBO = pitch * 128 and base-offset 0 (tests/test_gpu_umma_window.py).  The k*k
// taps thus share one activation load: 46 KB per 9 taps instead of 9 x 32 KB; only the weights still stream per tap
// (-42 % bytes landed per MMA at BN = 128 -- the forward kernel sat on its operand-feed floor).
// K-block order: (tap group sharing an activation view, channel chunk, tap).
#define MBW(bar, par) do { if (p.dbg & 32) mbar_wait((bar), (par)); else mbar_wait_spin((bar), (par)); } while (0)
template <int BN, bool HALO>
__global__ void __launch_bounds__(192, 1) tapconv_tc_kernel(const __grid_constant__ TcFwdParams p) {
  constexpr uint32_t kBBytes = BN * 128;
  constexpr uint32_t kStageBytes = HALO ? 2 * kBBytes : 2 * kABytes + 2 * kBBytes;  // HALO: a stage holds the weights only
  const uint32_t kAH = HALO ? p.a_tile_bytes : 0;                                  // one haloed activation tile (1 KB multiple)
  const uint32_t kAReg = 4 * kAH;                                                  // 2 stages x {hi, lo} in front of the ring
  constexpr uint32_t kIdesc = make_idesc(128, BN, 0, 0);
  constexpr uint32_t kIdescBf = make_idesc_bf16(128, BN);
  // TMEM layout (all 512 columns): [0,256) a ring of kAcc buffers for the MAIN term hi*hi -- the MMA warp
  // accumulates `chunk` K-blocks into one buffer, the epilogue promotes it to fp32 registers and frees it;
  // [256, 256+2*BN) two buffers (tile parity) for the CROSS terms hi*lo + lo*hi, which are 2^-11 smaller, so
  // their truncation drift is irrelevant and they stay in TMEM for the whole tile (read once at the end).
  // Promotion cost is bounded by the TMEM read rate (~64 B/clk): reading only the main block per chunk is what
  // makes a short chunk affordable.
  constexpr uint32_t kAcc = 256 / BN;
  constexpr uint32_t kCrossCol = 256;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem_a = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem = smem_a + kAReg;  // the stage ring
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint64_t* empty = full + kStages;
  uint64_t* tmem_full = empty + kStages;      // [kAcc]
  uint64_t* tmem_empty = tmem_full + kAcc;    // [kAcc]
  uint64_t* cross_empty = tmem_empty + kAcc;  // [2]
  uint64_t* a_full = cross_empty + 2;         // [2] (HALO)
  uint64_t* a_empty = a_full + 2;             // [2] (HALO)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_empty + 2);
  float* stat_sm = reinterpret_cast<float*>(smem + kStages * kStageBytes + 256);  // [2][4 warps][BN] (p.stats only)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = p.ntaps * p.kpt;
  const int kChunk = p.chunk;
  const int nchunks = (nkb + kChunk - 1) / kChunk;
  const int ntiles = p.ntiles;
  long long dbg_c0 = 0;
  unsigned long long dbg_t0 = 0;
  if ((p.dbg & 128) && threadIdx.x == 0) {  // experiment: effective SM clock during this launch
    dbg_c0 = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(dbg_t0));
  }

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, 1);
    }
    for (int i = 0; i < (int)kAcc; ++i) {
      mbar_init(tmem_full + i, 1);
      mbar_init(tmem_empty + i, 4);  // one arrive per epilogue warp
    }
    mbar_init(cross_empty + 0, 4);
    mbar_init(cross_empty + 1, 4);
    for (int i = 0; i < 2; ++i) {
      mbar_init(a_full + i, 1);
      mbar_init(a_empty + i, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // all 512 columns are allocated (one CTA per SM), so the allocation can only start at column 0 / lane 0: using the
  // literal keeps every TMEM address in uniform registers
  if (*tmem_slot != 0) __trap();
  constexpr uint32_t tmem_base = 0;
  const int tpg = p.tpg;  // HALO: taps per group (= taps sharing one activation view)

  if (warp == 0) {
    if (lane == 0) {
      prefetch_tmap(&p.b_hi);
      prefetch_tmap(&p.b_lo);
      uint32_t kbg = 0, acg = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const FwdTile t = fwd_decode<BN>(p, tile);
        for (int kb = 0; kb < nkb; ++kb, ++kbg) {
          if (HALO) {
            const int per_grp = p.kpt * tpg;
            const int grp = kb / per_grp, rem = kb - grp * per_grp;
            const int kc = rem / tpg, tig = rem - kc * tpg;
            const int ti = t.ph * p.ntaps + grp * tpg + tig;
            if (tig == 0) {  // a new (activation view, channel chunk): one haloed box each for hi and lo
              const uint32_t sa = acg & 1, ita = acg >> 1;
              if (ita > 0) MBW(a_empty + sa, (ita - 1) & 1);
              const int am = p.amap[ti];
              uint8_t* at = smem_a + sa * 2 * kAH;
              if (p.dbg & 2) {
                mbar_arrive(a_full + sa);
              } else {
                mbar_expect_tx(a_full + sa, 2 * p.a_box_bytes);
                tma_load_4d(at, &p.a_hi[am], a_full + sa, kc * 32, t.x0 - p.halo, t.y0 - p.halo, t.b0);
                tma_load_4d(at + kAH, &p.a_lo[am], a_full + sa, kc * 32, t.x0 - p.halo, t.y0 - p.halo, t.b0);
              }
              ++acg;
            }
            const uint32_t s = kbg % kStages, it = kbg / kStages;
            if (it > 0) MBW(empty + s, (it - 1) & 1);
            uint8_t* st = smem + s * kStageBytes;
            if (p.dbg & 2) {  // experiment: no data movement, only the barrier protocol
              mbar_arrive(full + s);
              continue;
            }
            mbar_expect_tx(full + s, kStageBytes);
            const int wrow = p.widx[ti] * p.Cout + t.n0;
            tma_load_2d(st, &p.b_hi, full + s, kc * 32, wrow);
            tma_load_2d(st + kBBytes, &p.b_lo, full + s, kc * 32, wrow);
            continue;
          }
          const uint32_t s = kbg % kStages, it = kbg / kStages;
          if (it > 0 && !(p.dbg & 512)) MBW(empty + s, (it - 1) & 1);  // dbg 512 (with 2): stages always "full"
          const int tap = kb / p.kpt, c0 = (kb - tap * p.kpt) * 32;
          const int ti = t.ph * p.ntaps + tap;
          const int am = p.amap[ti];
          uint8_t* st = smem + s * kStageBytes;
          if (p.dbg & 2) {  // experiment: no data movement, only the barrier protocol
            mbar_arrive(full + s);
            continue;
          }
          mbar_expect_tx(full + s, kStageBytes);
          // mixed mode: the "lo" maps view BF16 pair tensors, 64 elements (= the same 128 bytes) per 32-channel block
          const int c0l = p.mixed ? 2 * c0 : c0;
          tma_load_4d(st, &p.a_hi[am], full + s, c0, t.x0 + p.dx[ti], t.y0 + p.dy[ti], t.b0);
          tma_load_4d(st + kABytes, &p.a_lo[am], full + s, c0l, t.x0 + p.dx[ti], t.y0 + p.dy[ti], t.b0);
          const int wrow = p.widx[ti] * p.Cout + t.n0;
          tma_load_2d(st + 2 * kABytes, &p.b_hi, full + s, c0, wrow);
          tma_load_2d(st + 2 * kABytes + kBBytes, &p.b_lo, full + s, c0l, wrow);
        }
      }
    }
  } else if (warp == 1) {
    // MMA issue.  The WHOLE warp runs the (warp-uniform) control flow and the barrier waits; one elected lane issues the
    // tcgen05 instructions.  Under `if (lane == 0)` the compiler could not prove the operands uniform and wrapped every
    // UTCHMMA in an ELECT / R2UR.BROADCAST / BRA.U.ANY loop: ~14 instructions and ~107 clocks of issue per 64-clock MMA
    // (measured with the data movement switched off) -- the tensor pipe idled 40 % of the time waiting for its one thread.
    {
      uint32_t kbg = 0, cg = 0, tl = 0, acm = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tl) {
        const FwdTile t = fwd_decode<BN>(p, tile);
        const uint32_t tcross = tmem_base + kCrossCol + (tl & 1) * BN;
        if (tl >= 2) MBW(cross_empty + (tl & 1), ((tl >> 1) - 1) & 1);  // epilogue has read tile tl-2's cross block
        for (int ch = 0; ch < nchunks; ++ch, ++cg) {
          const uint32_t buf = cg % kAcc, use = cg / kAcc;
          if (use > 0) MBW(tmem_empty + buf, (use - 1) & 1);
          tc_fence_after();
          const uint32_t tacc = tmem_base + buf * BN;
          const int nk = min(kChunk, nkb - ch * kChunk);
          for (int j = 0; j < nk; ++j, ++kbg) {
            const uint32_t s = kbg % kStages, it = kbg / kStages;
            const uint32_t sa = smem_u32(smem + ((p.dbg & 16) ? 0 : s) * kStageBytes);  // dbg 16: constant operand addresses
            uint64_t a_hi, a_lo, b_hi, b_lo;
            bool last_of_view = false;
            uint32_t sav = 0;
            if (HALO) {
              const int kb = ch * kChunk + j;
              const int per_grp = p.kpt * tpg;
              const int grp = kb / per_grp, rem = kb - grp * per_grp;
              const int tig = rem % tpg;
              const int ti = t.ph * p.ntaps + grp * tpg + tig;
              sav = acm & 1;
              if (tig == 0) MBW(a_full + sav, (acm >> 1) & 1);
              last_of_view = tig == tpg - 1;
              const uint32_t pitch = 8 + 2 * p.halo;
              uint32_t off = (uint32_t)((p.dy[ti] + p.halo) * (int)pitch + p.dx[ti] + p.halo) * 128;
              if (p.dbg & 4) off = 0;  // experiment (wrong results): every window 1024-aligned
              const uint32_t ab = smem_u32(smem_a + sav * 2 * kAH) + off;
              a_hi = make_desc(ab, 16, pitch * 128);  // shifted window: 8-row groups one image row apart
              a_lo = make_desc(ab + kAH, 16, pitch * 128);
              b_hi = make_desc(sa, 16, 1024);
              b_lo = make_desc(sa + kBBytes, 16, 1024);
            } else {
              a_hi = make_desc(sa, 16, 1024);
              a_lo = make_desc(sa + kABytes, 16, 1024);
              b_hi = make_desc(sa + 2 * kABytes, 16, 1024);
              b_lo = make_desc(sa + 2 * kABytes + kBBytes, 16, 1024);
            }
            MBW(full + s, it & 1);
            // (no tcgen05 fence here: the stage was filled by TMA through the mbarrier's complete_tx, which the wait
            // acquires; a fence::after_thread_sync per K block is only needed where TMEM changes hands, see above)
            if (p.dbg & 256) tc_fence_after();
            if (!elect_one()) {
            } else if (p.dbg & 1) {  // (dbg bit 0: experiment without MMAs)
            } else if (!p.mixed && (p.dbg & 64)) {  // (round-1 order: the accumulator alternates with every instruction)
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint64_t ko = (uint64_t)(k * 2);  // +32 bytes (8 fp32 of K) in the 16B-unit start-address field
                umma_tf32(tacc, a_hi + ko, b_hi + ko, kIdesc, (j | k) != 0);        // main term -> ring buffer
                umma_tf32(tcross, a_hi + ko, b_lo + ko, kIdesc, (ch | j | k) != 0);  // cross terms -> per-tile block
                umma_tf32(tcross, a_lo + ko, b_hi + ko, kIdesc, 1);
              }
            } else if (!p.mixed) {
              // the 4 K slices of one product back to back: the accumulator changes twice per K block instead of 8 times
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_tf32(tacc, a_hi + (uint64_t)(k * 2), b_hi + (uint64_t)(k * 2), kIdesc, (j | k) != 0);
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_tf32(tcross, a_hi + (uint64_t)(k * 2), b_lo + (uint64_t)(k * 2), kIdesc, (ch | j | k) != 0);
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_tf32(tcross, a_lo + (uint64_t)(k * 2), b_hi + (uint64_t)(k * 2), kIdesc, 1);
            } else {
              // main term in TF32 (exact products); the two cross terms are ~2^-12 of the result, so BF16 inputs
              // (rel. 2^-9) keep them to ~2^-20: kind::f16 runs at twice the TF32 rate => 4 + 2 + 2 half-cost MMAs.
              // In the pair tile a 128-byte row is [32 x bf16(hi) | 32 x bf16(lo)]: +0 B / +64 B pick the half,
              // +32 B steps the 16-element K slice.
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_tf32(tacc, a_hi + (uint64_t)(k * 2), b_hi + (uint64_t)(k * 2), kIdesc, (j | k) != 0);
#pragma unroll
              for (int k = 0; k < 2; ++k) {
                const uint64_t ko = (uint64_t)(k * 2);
                umma_bf16(tcross, a_lo + ko, b_lo + 4 + ko, kIdescBf, (ch | j | k) != 0);  // bf16(a_hi) . bf16(b_lo)
                umma_bf16(tcross, a_lo + 4 + ko, b_lo + ko, kIdescBf, 1);                  // bf16(a_lo) . bf16(b_hi)
              }
            }
            if (elect_one()) umma_commit(empty + s);
            if (HALO && last_of_view) {  // every tap of this activation view has been issued: its tile may be refilled
              if (elect_one()) umma_commit(a_empty + sav);
              ++acm;
            }
            __syncwarp();
          }
          if (elect_one()) umma_commit(tmem_full + buf);
          __syncwarp();
        }
      }
    }
  } else {
    // ---- epilogue: 4 warps, warp%4 selects the TMEM lane quadrant ----
    const int q = warp & 3;
    const int m = q * 32 + lane;  // accumulator row == tile pixel
    const int xi = m % p.bw, yi = (m / p.bw) % p.bh, bi = m / (p.bw * p.bh);
    uint32_t cg = 0, tl = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tl) {
      const FwdTile t = fwd_decode<BN>(p, tile);
      float acc[BN];
#pragma unroll
      for (int i = 0; i < BN; ++i) acc[i] = 0.f;
      for (int ch = 0; ch < nchunks; ++ch, ++cg) {
        const uint32_t buf = cg % kAcc, use = cg / kAcc;
        mbar_wait(tmem_full + buf, use & 1);
        tc_fence_after();
        if (!(p.dbg & 8)) {  // dbg 8: experiment without the promotion's TMEM reads
#pragma unroll
          for (int j = 0; j < BN / 32; j += 2) {
            uint32_t va[32], vb[32];
            const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + buf * BN + (uint32_t)(j * 32);
            tmem_ld_32x32_x2(ta, ta + 32, va, vb);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              acc[j * 32 + i] += __uint_as_float(va[i]);
              acc[j * 32 + 32 + i] += __uint_as_float(vb[i]);
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tmem_empty + buf);
      }
      {  // the last chunk's commit also covers every cross-term MMA of this tile: add the cross block once
#pragma unroll
        for (int j = 0; j < BN / 32; j += 2) {
          uint32_t va[32], vb[32];
          const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + kCrossCol + (tl & 1) * BN + (uint32_t)(j * 32);
          tmem_ld_32x32_x2(ta, ta + 32, va, vb);
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            acc[j * 32 + i] += __uint_as_float(va[i]);
            acc[j * 32 + 32 + i] += __uint_as_float(vb[i]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(cross_empty + (tl & 1));
      }
      const int b = t.b0 + bi;
      if (p.bias) {
#pragma unroll
        for (int i = 0; i < BN; ++i) acc[i] += p.bias[t.n0 + i];
      }
      if (p.stats) {
        // BatchNorm statistics from the convolution epilogue (nn.SpatialBatchNormalization, models.lua:65,70): per
        // tile the column sums of z and z^2 over its (valid) 128 pixels.  Within a warp a transpose-reduce (31
        // shuffles per 32 columns) leaves lane l with the total of column l; the 4 warps meet in shared memory and
        // one thread per column writes the tile's partial.  Every sum has a fixed order => replicas stay identical.
        const bool valid = b < p.B;
#pragma unroll
        for (int j = 0; j < BN / 32; ++j) {
          float x[32], y[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float v = valid ? acc[j * 32 + i] : 0.f;
            x[i] = v;
            y[i] = v * v;
          }
#pragma unroll
          for (int sft = 16; sft >= 1; sft >>= 1) {
            const bool up = (lane & sft) != 0;
#pragma unroll
            for (int k = 0; k < sft; ++k) {
              const float sx = up ? x[k] : x[k + sft], kx = up ? x[k + sft] : x[k];
              const float sy = up ? y[k] : y[k + sft], ky = up ? y[k + sft] : y[k];
              x[k] = kx + __shfl_xor_sync(0xffffffffu, sx, sft);
              y[k] = ky + __shfl_xor_sync(0xffffffffu, sy, sft);
            }
          }
          stat_sm[(0 * 4 + q) * BN + j * 32 + lane] = x[0];
          stat_sm[(1 * 4 + q) * BN + j * 32 + lane] = y[0];
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const int et = (int)threadIdx.x - 64;  // 0..127 over the 4 epilogue warps
        if (et < BN) {
          const float s0 = (stat_sm[0 * BN + et] + stat_sm[1 * BN + et]) + (stat_sm[2 * BN + et] + stat_sm[3 * BN + et]);
          const float s1 = (stat_sm[4 * BN + et] + stat_sm[5 * BN + et]) + (stat_sm[6 * BN + et] + stat_sm[7 * BN + et]);
          const int mt = tile / (p.Cout / BN);
          p.stats[((int64_t)mt * 2 + 0) * p.Cout + t.n0 + et] = s0;
          p.stats[((int64_t)mt * 2 + 1) * p.Cout + t.n0 + et] = s1;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
      if (b < p.B) {
        const int Y = p.out_scale * (t.y0 + yi) + (t.ph >> 1) * (p.out_scale - 1);
        const int X = p.out_scale * (t.x0 + xi) + (t.ph & 1) * (p.out_scale - 1);
        float* orow = p.out + (((int64_t)b * p.out_H + Y) * p.out_W + X) * p.Cout + t.n0;
#pragma unroll
        for (int i = 0; i < BN; i += 4)
          *reinterpret_cast<float4*>(orow + i) = make_float4(acc[i], acc[i + 1], acc[i + 2], acc[i + 3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
  if ((p.dbg & 128) && threadIdx.x == 0 && blockIdx.x == 0) {
    unsigned long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    printf("tapconv<%d,%d> block 0: %lld cycles in %llu ns = %.3f GHz, %d tiles x %d kblocks\n", BN, (int)HALO, clock64() - dbg_c0,
           t1 - dbg_t0, (double)(clock64() - dbg_c0) / (double)(t1 - dbg_t0), (ntiles + (int)gridDim.x - 1) / (int)gridDim.x, nkb);
  }
}

#undef MBW
// ------------------------------------------------------------------------------------------------
// wgrad: D[n (M=128 of Cout)][c (BN of Cin)] += sum over a pixel range of dY[p][n] * X[p+off][c]
// grid: x = tile-tap, y = mtile * ntiles_n + ntile, z = K split.  Output accumulated with atomics.
// ------------------------------------------------------------------------------------------------
template <int BN>
__global__ void __launch_bounds__(192, 1) wgrad_tc_kernel(const __grid_constant__ TcWgParams p) {
  constexpr uint32_t kBox = 32 * 128;  // one (32 ch x 32 px) box = 4 KB
  constexpr uint32_t kAB = 4 * kBox;   // M = 128 channels of dY
  constexpr uint32_t kBB = (BN / 32) * kBox;
  constexpr uint32_t kStageBytes = 2 * kAB + 2 * kBB;
  constexpr uint32_t kIdesc = make_idesc(128, BN, 1, 1);
  constexpr uint32_t kCrossCol = 256;  // main ring: 2 buffers at [0, 2*BN); cross terms: [256, 256+BN) for the whole tile
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint64_t* empty = full + kStages;
  uint64_t* tmem_full = empty + kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tt = blockIdx.x;
  const int ntn = p.Cin / BN;
  const int m0 = (blockIdx.y / ntn) * 128, c0 = (blockIdx.y % ntn) * BN;
  const int kb_begin = blockIdx.z * p.kb_per_split;
  const int kb_end = min(p.kblocks, kb_begin + p.kb_per_split);
  const int nkb = kb_end - kb_begin;
  const int kChunk = p.chunk;
  const int nchunks = (nkb + kChunk - 1) / kChunk;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(tmem_full + i, 1);
      mbar_init(tmem_empty + i, 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (*tmem_slot != 0) __trap();  // the whole TMEM is allocated: base 0 (keeps TMEM addresses uniform, see tapconv)
  constexpr uint32_t tmem_base = 0;

  if (nkb > 0) {
    if (warp == 0) {
      if (lane == 0) {
        const int ph = p.phase[tt], dyo = p.dy[tt], dxo = p.dx[tt];
        for (int i = 0; i < nkb; ++i) {
          const int s = i % kStages, it = i / kStages;
          if (it > 0) mbar_wait(empty + s, (it - 1) & 1);
          const int kb = kb_begin + i;
          int b0, y0, x0;
          if (p.bb == 1) {
            const int per_img = p.tiles_x * p.tiles_y;
            b0 = kb / per_img;
            const int r = kb % per_img;
            y0 = (r / p.tiles_x) * p.bh;
            x0 = (r % p.tiles_x) * p.bw;
          } else {
            b0 = kb * p.bb;
            y0 = 0;
            x0 = 0;
          }
          uint8_t* st = smem + s * kStageBytes;
          mbar_expect_tx(full + s, kStageBytes);
          // 5-D maps (32 ch, w, h, b, channel-group): ONE bulk copy lands [group][pixel][32 ch] = all the 4 KB
          // boxes of an operand (16 single-box copies per stage made the kernel TMA-issue bound)
          tma_load_5d(st, &p.dy_hi[ph], full + s, 0, x0, y0, b0, m0 / 32);
          tma_load_5d(st + kAB, &p.dy_lo[ph], full + s, 0, x0, y0, b0, m0 / 32);
          tma_load_5d(st + 2 * kAB, &p.x_hi, full + s, 0, x0 + dxo, y0 + dyo, b0, c0 / 32);
          tma_load_5d(st + 2 * kAB + kBB, &p.x_lo, full + s, 0, x0 + dxo, y0 + dyo, b0, c0 / 32);
        }
      }
    } else if (warp == 1) {
      {  // whole warp, one elected lane issues (see tapconv_tc_kernel)
        int i = 0;
        for (int ch = 0; ch < nchunks; ++ch) {
          const uint32_t buf = ch & 1, use = ch >> 1;
          if (use > 0) mbar_wait(tmem_empty + buf, (use - 1) & 1);
          tc_fence_after();
          const uint32_t tacc = tmem_base + buf * BN, tcross = tmem_base + kCrossCol;
          const int nk = min(kChunk, nkb - ch * kChunk);
          for (int j = 0; j < nk; ++j, ++i) {
            const int s = i % kStages, it = i / kStages;
            mbar_wait(full + s, it & 1);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + s * kStageBytes);
            // MN-major operands.  layout 1 = SWIZZLE_128B_BASE32B: the only smem layout tcgen05 accepts for
            // MN-major tf32 (4 pixel rows x 128 B per swizzle atom, 32 B chunks XOR row%4; TMA side
            // SWIZZLE_128B_ATOM_32B).  LBO = distance between 32-channel groups (one 4 KB box),
            // SBO = distance between 4-pixel groups (512 B).
            const uint64_t a_hi = make_desc(sa, kBox, 512, 1), a_lo = make_desc(sa + kAB, kBox, 512, 1);
            const uint64_t b_hi = make_desc(sa + 2 * kAB, kBox, 512, 1), b_lo = make_desc(sa + 2 * kAB + kBB, kBox, 512, 1);
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint64_t ko = (uint64_t)(k * 64);  // +1024 bytes = next 8 pixels
                umma_tf32(tacc, a_hi + ko, b_hi + ko, kIdesc, (j | k) != 0);        // main term -> ring buffer
                umma_tf32(tcross, a_hi + ko, b_lo + ko, kIdesc, (ch | j | k) != 0);  // cross terms stay in TMEM
                umma_tf32(tcross, a_lo + ko, b_hi + ko, kIdesc, 1);
              }
              umma_commit(empty + s);
            }
            __syncwarp();
          }
          if (elect_one()) umma_commit(tmem_full + buf);
          __syncwarp();
        }
      }
    } else {
      const int q = warp & 3;
      const int n = m0 + q * 32 + lane;
      float acc[BN];
#pragma unroll
      for (int i = 0; i < BN; ++i) acc[i] = 0.f;
      for (int ch = 0; ch < nchunks; ++ch) {
        const uint32_t buf = ch & 1, use = ch >> 1;
        mbar_wait(tmem_full + buf, use & 1);
        tc_fence_after();
#pragma unroll
        for (int j = 0; j < BN / 32; j += 2) {
          uint32_t va[32], vb[32];
          const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + buf * BN + (uint32_t)(j * 32);
          tmem_ld_32x32_x2(ta, ta + 32, va, vb);
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            acc[j * 32 + i] += __uint_as_float(va[i]);
            acc[j * 32 + 32 + i] += __uint_as_float(vb[i]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tmem_empty + buf);
      }
#pragma unroll
      for (int j = 0; j < BN / 32; j += 2) {  // cross terms: complete once the last chunk has been committed
        uint32_t va[32], vb[32];
        const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + kCrossCol + (uint32_t)(j * 32);
        tmem_ld_32x32_x2(ta, ta + 32, va, vb);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          acc[j * 32 + i] += __uint_as_float(va[i]);
          acc[j * 32 + 32 + i] += __uint_as_float(vb[i]);
        }
      }
      float* orow = p.out + ((int64_t)tt * p.Cout + n) * p.Cin + c0;
#pragma unroll
      for (int i = 0; i < BN; ++i) atomicAdd(orow + i, acc[i]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}
