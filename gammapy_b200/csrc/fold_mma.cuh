// fold_mma.cuh -- K3 + K4 fused, second generation: the fold + Cash kernel as a warp-independent pipeline.
//
//   apply_exposure -> apply_edisp -> stack -> + background -> clip -> masked Cash
//   (datasets/evaluator.py:346-377, datasets/utils.py:66-84, maps/wcs/ndmap.py:1138-1181,
//    datasets/map.py:775-822, stats/fit_statistics.py:281-293, fit_statistics_cython.pyx:55-85)
//
// Design:
//   * One persistent CTA per SM: kMmaConsumers consumer warps + 1 producer warp, no CTA barrier after start-up.
//   * The unit of work is a SUB-ITEM: 32 consecutive pixels of one (tile, parameter vector) work item, all energies.
//     The producer warp claims tiles (statically or from a global counter), and for every sub-item fills one
//     STAGE of a shared-memory ring with 2-D tensor-map TMA loads (cp.async.bulk.tensor, SWIZZLE_128B):
//     exposure [nT][32 px] f32, background / counts [nR][32 px] f32, mask [nR][32 px] u8, plus the tile's work
//     descriptor and a header.  full[] / empty[] mbarriers per stage; consumer warps take stages in sequence order
//     from a shared-memory counter, so a heavy sub-item (many sources) never stalls the others.
//     Sub-items no source reaches do not fetch the exposure at all (40 % of the C3 tiles).
//   * A consumer warp does the whole fold of its 32 pixels in registers, using the fp64 tensor-core instruction
//     DMMA.8x8x4 (mma.sync.m8n8k4.f64; full fp64 rate on B200, profiles/micro/dmma_bench_b200.log) for the
//     true -> reco contraction:
//         A[p][t] = v[p][t] = exposure[t,p] * sum_s F_s[t] * C_s[t,p]     (thread (g, tg) owns pixel quad 4g..4g+3
//                                                                         x true bin t(tg): exactly what it loaded)
//         B[t][r] = 1e4 * pdf[t][r]                                        (fragment-ordered table, non-zero band only)
//         C[p][r] = npred_signal                                           (32 fp64 accumulators per thread)
//     so v never travels through shared memory and there is nothing to synchronise.  The 128-byte swizzle makes the
//     exposure / background / counts reads (row = t or r, 16-byte chunk = g) bank-conflict free: a k-step takes the
//     rows {8 blk + 2 tg + h} (h = k-step parity) so that the four rows of a quarter warp differ in bits 1-2.
//   * The per-source PSF-convolved cubes are read with 128-bit ld.global.nc (L1 no-allocate, L2 evict-last) through
//     per-thread pointers (a zero word with stride 0 outside the cutout), prefetched 4 (2) k-steps ahead in
//     registers.  More than four sources on a tile, or a second edisp group, are further passes over the k-steps
//     that accumulate into the same C registers (the contraction is linear).
//   * Epilogue per thread: its 4 pixels x 8 reco bins; background model, clip, Cash term; warp-shuffle sum; one
//     partial per sub-item (deterministic: fixed order inside the warp, reduce_stat_kernel adds the partials in order).
#pragma once
#include "kernels.cuh"

namespace gpy {

constexpr int kMmaSub = 32;                            // pixels per sub-item
constexpr int kMmaSubs = kFoldTPX / kMmaSub;           // sub-items per tile
constexpr int kMmaMaxNT = 64, kMmaMaxNR = 32;          // axes this kernel handles (larger: the staged / generic kernels)
constexpr int kMmaKSteps = kMmaMaxNT / 4;
constexpr int kMmaNTiles = kMmaMaxNR / 8;
// warps per CTA: 11 -> 352 threads x 168 registers, 8 -> 256 threads x 255 registers (ptxas sizes the register budget
// for thread counts rounded up to 128)

constexpr int kLogTabBytes = 128 * 16;                  // fast_log table: 128 x {1 / m_hi, log(m_hi)}

// Edisp matrices of one dataset in DMMA B-fragment order, non-zero band only (built by the host, see
// build_mma_table).  k-step ks covers the true bins t = 8 (ks >> 1) + (ks & 1) + 2 tg, n-tile n the reco bins 8 n + g.
struct MmaTableHead {
  unsigned long long bits[2];  // per edisp group: 4 bits per k-step = the n-tiles with a non-zero coefficient
  int ks_lo[2], ks_hi[2];      // active k-step range
  int n_groups, pad_[7];
  // double frag[n_groups][kMmaKSteps][kMmaNTiles][32] follows (byte 64): dense, zero where bits says inactive
};
constexpr int kMmaGroupDoubles = kMmaKSteps * kMmaNTiles * 32;
static_assert(sizeof(MmaTableHead) == 64, "MmaTableHead");

struct MmaParams {
  unsigned lay_tab, lay_log, lay_bars;  // shared-memory offsets behind the per-warp areas: cached table, log table, mbarriers
  int tab_bytes;                 // bytes of the table cached in shared memory (dataset tab_ds)
  int tab_ds;                    // dataset whose table is cached (-1: none)
  const double2* logtab;         // fast_log table (device)
  int debug_skip;                // GPY_MMA_SKIP (profiling only): 1 skips the k-loops, 2 the epilogue, 4 the exposure loads
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tma_prefetch_2d(const void* tmap, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(tmap), "r"(c0), "r"(c1) : "memory");
}

__device__ __forceinline__ float4 ldg_f4_stream(const float* p, uint64_t pol) {
  float4 v;
  asm("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
      : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
      : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ unsigned ldg_u32_stream(const void* p, uint64_t pol) {
  unsigned v;
  asm("ld.global.nc.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
  return v;
}

// D += A * B on the fp64 tensor core path (DMMA.8x8x4).
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
      : "+d"(d0), "+d"(d1)
      : "d"(a), "d"(b));
}

// log(x) for finite x > 1e-25 (normal): x = 2^e m, m in [1, 2); m = m_hi (1 + r) with m_hi from the top 7 mantissa
// bits, |r| < 2^-8; log(x) = e ln2 + log(m_hi) + log1p(r), log1p by a degree-7 polynomial (|error| < 1e-17 absolute).
// Straight-line code: four calls interleave (the libdevice log has data-dependent branches).
__device__ __forceinline__ double fast_log(double x, const double2* __restrict__ tab) {
  const long long ix = __double_as_longlong(x);
  const int e = (int)(ix >> 52) - 1023;
  const double2 t = tab[(int)(ix >> 45) & 127];
  const double m = __longlong_as_double((ix & 0x000fffffffffffffLL) | 0x3ff0000000000000LL);
  const double r = fma(m, t.x, -1.0);
  const double r2 = r * r;
  const double a = fma(r, 1.0 / 3.0, -0.5), b = fma(r, 0.2, -0.25), c = fma(r, 1.0 / 7.0, -1.0 / 6.0);
  double q = fma(r2, c, b);
  q = fma(r2, q, a);
  const double p = fma(r2, q, r);
  return fma((double)e, 0.6931471805599453, t.y) + p;
}

struct MmaSrc {  // per-thread view of up to four sources of one pass
  const float* cptr[4];  // this thread's pixel quad in plane 0 of the source's cube; a zero word when outside the cutout
  int pstr[4];           // plane stride in floats (0 outside the cutout / no source)
};

// One pass over the k-steps [k_begin, k_end) (k_begin a multiple of UN) for NL sources:
//   acc[j][n][i] += sum_t v_j[t] * B[t][8 n + 2 tg + i],   v_j[t] = exposure[t, p_j] * sum_l F_l[t] * C_l[t, p_j].
// Everything a k-step reads sits at a compile-time offset from a per-lane pointer that advances once per loop
// iteration (UN k-steps): the exposure tile (swizzled; two lane constants), the zero-padded flux rows fs[l][64] and the
// dense fragment table tabg[ks][n][32] in shared memory, the cube rows t, t+1, t+8, t+9 (UN = 4) through 32-bit byte
// steps.  The cube loads of the next iteration are issued before the DMMAs of this one (PD = UN k-steps ahead).
template <int NL, int UN>
__device__ __forceinline__ void fold_pass(double (&acc)[4][kMmaNTiles][2], const MmaSrc& S, const double* fs,
                                          const unsigned char* es, const double* tabg, unsigned long long bits,
                                          int k_begin, int k_end, int nT, int g, int tg, int lane, uint64_t pol) {
  static_assert(UN == 2 || UN == 4, "k-steps per iteration");
  const int t_lane = 2 * tg;
  // per-lane constants
  const unsigned char* ep = es + t_lane * 128 + k_begin * 512;        // row 8 blk + 2 tg (+ h), blk = ks >> 1
  const int ex0 = (g ^ (2 * tg)) << 4, ex1 = ((g ^ (2 * tg + 1)) << 4) + 128;
  const double* fp = fs + t_lane + k_begin * 4;
  const double* tp = tabg + (size_t)k_begin * (kMmaNTiles * 32) + lane;
  const char* cp[NL];
  int s1[NL];
#pragma unroll
  for (int l = 0; l < NL; ++l) {
    s1[l] = S.pstr[l] * 4;
    cp[l] = reinterpret_cast<const char*>(S.cptr[l]) + (size_t)(4 * k_begin + t_lane) * (size_t)s1[l];
  }
  int t0 = 4 * k_begin + t_lane;  // this lane's first row of the iteration
  unsigned long long mb = bits >> (4 * k_begin);
  float4 cbuf[UN][NL];
  auto issue = [&](float4 (&dst)[UN][NL], int tbase) {
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int dt = (u >> 1) * 8 + (u & 1);
      if (tbase + dt < nT) {  // rows past the axis keep their (finite) old value; their flux is zero
#pragma unroll
        for (int l = 0; l < NL; ++l)
          dst[u][l] = ldg_f4_stream(reinterpret_cast<const float*>(cp[l] + (size_t)(unsigned)(dt * s1[l])), pol);
      }
    }
  };
#pragma unroll
  for (int u = 0; u < UN; ++u)
#pragma unroll
    for (int l = 0; l < NL; ++l) cbuf[u][l] = make_float4(0.f, 0.f, 0.f, 0.f);
  issue(cbuf, t0);
  for (int k0 = k_begin; k0 < k_end; k0 += UN) {
    float4 c[UN][NL];
#pragma unroll
    for (int u = 0; u < UN; ++u)
#pragma unroll
      for (int l = 0; l < NL; ++l) c[u][l] = cbuf[u][l];
    // next iteration's cube rows
#pragma unroll
    for (int l = 0; l < NL; ++l) cp[l] += (size_t)(unsigned)(4 * UN * s1[l]);
    if (k0 + UN < k_end) issue(cbuf, t0 + 4 * UN);
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const unsigned m = (unsigned)(mb >> (4 * u)) & 15u;
      const int blk = u >> 1, h = u & 1;
      const float4 e = *reinterpret_cast<const float4*>(ep + blk * 1024 + (h ? ex1 : ex0));
      double bfr[kMmaNTiles];
#pragma unroll
      for (int n = 0; n < kMmaNTiles; ++n) bfr[n] = tp[(u * kMmaNTiles + n) * 32];
      double F[NL];
#pragma unroll
      for (int l = 0; l < NL; ++l) F[l] = fp[l * kMmaMaxNT + blk * 8 + h];
      double v0 = F[0] * (double)c[u][0].x, v1 = F[0] * (double)c[u][0].y, v2 = F[0] * (double)c[u][0].z,
             v3 = F[0] * (double)c[u][0].w;
#pragma unroll
      for (int l = 1; l < NL; ++l) {
        v0 = fma(F[l], (double)c[u][l].x, v0);
        v1 = fma(F[l], (double)c[u][l].y, v1);
        v2 = fma(F[l], (double)c[u][l].z, v2);
        v3 = fma(F[l], (double)c[u][l].w, v3);
      }
      v0 *= (double)e.x;
      v1 *= (double)e.y;
      v2 *= (double)e.z;
      v3 *= (double)e.w;
#pragma unroll
      for (int n = 0; n < kMmaNTiles; ++n) {
        if (m & (1u << n)) {  // warp uniform
          dmma884(acc[0][n][0], acc[0][n][1], v0, bfr[n]);
          dmma884(acc[1][n][0], acc[1][n][1], v1, bfr[n]);
          dmma884(acc[2][n][0], acc[2][n][1], v2, bfr[n]);
          dmma884(acc[3][n][0], acc[3][n][1], v3, bfr[n]);
        }
      }
    }
    ep += UN * 512;
    fp += UN * 4;
    tp += UN * kMmaNTiles * 32;
    t0 += 4 * UN;
    mb >>= 4 * UN;
  }
}

struct EpiArgs {
  const float* bkgp;
  const float* cntp;
  const unsigned char* mskp;
  const double* bf;        // background factors of this (dataset, vector)
  double* npred_out;
  const double2* logtab;
  double log_trunc;
  uint64_t pol;
  int Pn, col;             // pixels per plane, first pixel of this thread's quad
  int nR, tg, flags, include_background;
  bool live;               // false for a quad beyond the last pixel (col then points at a valid pixel)
};

// Background model, clip and Cash terms of the 8 reco rows x 4 pixels a thread holds (fit_statistics_cython.pyx:55-85:
// mu' - n log mu' with mu' = mu > 1e-25 ? mu : 1e-25, the log only where n > 0).  Returns the thread's sum.
// FAST: fast_log, interleaved over the four pixels and only when some lane of the warp has a count in the row; the
// caller falls back to FAST = false (libdevice log) when the warp's sum is not finite.  The clip at 0 (map.py:787)
// only matters to the npred cube: the Cash term of a negative mu is that of 1e-25 either way.
template <bool WRITE_NPRED, bool FAST>
__device__ __forceinline__ double mma_epilogue(const double (&acc)[4][kMmaNTiles][2], const EpiArgs& A) {
  double stat = 0.0;
  const bool has_bkg = (A.flags & 2) != 0, has_counts = (A.flags & 4) != 0, has_mask = (A.flags & 8) != 0;
  // Two rounds of two n-tiles (four reco rows) each; the round is a real loop (code size: the instruction cache holds
  // the whole hot path), its accumulators are selected into a 16-value window.
#pragma unroll 1
  for (int half = 0; half < kMmaNTiles / 2; ++half) {
    double win[4][2][2];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int nn = 0; nn < 2; ++nn)
#pragma unroll
        for (int i = 0; i < 2; ++i) win[j][nn][i] = half ? acc[j][2 + nn][i] : acc[j][nn][i];
    float4 bg[4], cn[4];
    unsigned mk[4];
    double fac[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int r = 16 * half + 8 * (q >> 1) + 2 * A.tg + (q & 1);
      const int rc = min(r, A.nR - 1);  // rows past the axis read a valid row and are masked out
      const size_t off = (size_t)rc * A.Pn + A.col;
      bg[q] = make_float4(0.f, 0.f, 0.f, 0.f);
      cn[q] = make_float4(0.f, 0.f, 0.f, 0.f);
      mk[q] = 0x01010101u;
      fac[q] = 0.0;
      if (has_bkg) {
        bg[q] = ldg_f4_stream(A.bkgp + off, A.pol);
        fac[q] = __ldg(A.bf + rc);
      }
      if (has_counts) {
        cn[q] = ldg_f4_stream(A.cntp + off, A.pol);
        if (has_mask) mk[q] = ldg_u32_stream(A.mskp + off, A.pol);
      }
      if (r >= A.nR || !A.live) mk[q] = 0u;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int nn = q >> 1, i = q & 1;
      double mu[4] = {win[0][nn][i], win[1][nn][i], win[2][nn][i], win[3][nn][i]};
      if (A.include_background) {
        mu[0] += (double)bg[q].x * fac[q];
        mu[1] += (double)bg[q].y * fac[q];
        mu[2] += (double)bg[q].z * fac[q];
        mu[3] += (double)bg[q].w * fac[q];
      }
      if (WRITE_NPRED) {
        const int r = 16 * half + 8 * nn + 2 * A.tg + i;
        if (r < A.nR && A.live) {
          if (A.include_background) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (mu[j] < 0.0) mu[j] = 0.0;  // npred_total.data[npred_total.data < 0.0] = 0 (map.py:787)
          }
          double* o = A.npred_out + (size_t)r * A.Pn + A.col;
          *reinterpret_cast<double2*>(o) = make_double2(mu[0], mu[1]);
          *reinterpret_cast<double2*>(o + 2) = make_double2(mu[2], mu[3]);
        }
      } else {
        const float nf[4] = {cn[q].x, cn[q].y, cn[q].z, cn[q].w};
        bool big[4];
        bool need = false;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          big[j] = mu[j] > kTrunc;
          need |= nf[j] > 0.f;
        }
        need = need && mk[q] != 0u;
        double lg[4] = {0.0, 0.0, 0.0, 0.0};
        if (__any_sync(0xffffffffu, need)) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const double l = FAST ? fast_log(big[j] ? mu[j] : 1.0, A.logtab) : log(big[j] ? mu[j] : 1.0);
            lg[j] = big[j] ? l : A.log_trunc;
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          // n = 0: the product vanishes (lg is finite or, with FAST = false, the reference's own non-finite value
          // times a count that the reference also multiplies only when n > 0)
          double term = big[j] ? mu[j] : kTrunc;
          if (FAST)
            term = fma(-(double)nf[j], lg[j], term);
          else if (nf[j] > 0.f)
            term -= (double)nf[j] * lg[j];
          if ((mk[q] >> (8 * j)) & 0xffu) stat += term;
        }
      }
    }
  }
  return stat;
}

enum : int {
  kHdrHasSrc = 1, kHdrHasBkg = 2, kHdrHasCounts = 4, kHdrHasMask = 8,
};

// Per-tile information a warp keeps in shared memory (two slots: the tile it works on and the next one).
struct MmaTileInfo {
  int w, tile, b, d;            // work position (descriptor slot), tile, parameter vector, dataset
  int evalidx, nsrc, src_mask, pt;  // (dataset, vector) index; overlapping sources; bit 4 sub + l: source l reaches
                                    // sub-item sub; first pixel of the tile in its dataset
  int nx, nT, nR, Pn;
  int G, dflags, pad0_, pad1_;
};
static_assert(sizeof(MmaTileInfo) == 64, "MmaTileInfo");

// per-warp shared memory: two exposure buffers, two descriptors, two tile infos, the flux rows of a pass
constexpr int kWpExp = 0;                              // [2][64][128 B], SWIZZLE_128B
constexpr int kWpDesc = 2 * kMmaMaxNT * 128;           // [2][448 B]
constexpr int kWpInfo = kWpDesc + 2 * 448;             // [2] MmaTileInfo
constexpr int kWpFs = kWpInfo + 2 * 64;                // [2][64] doubles
constexpr int kWarpBytes = kWpFs + 2 * kMmaMaxNT * 8;
static_assert(kWarpBytes == 18432, "per-warp layout");

// One CTA per SM, NC self-sufficient warps.  A warp claims tiles (128 pixels x one parameter vector) from a global
// counter two tiles ahead, and walks the four 32-pixel sub-items of each: while it folds one sub-item it has the
// exposure tile of the next one in flight (its own 2-D tensor-map TMA load into the other of its two buffers) and the
// background / counts / mask boxes of the next one on their way to L2 (TMA prefetch).  No warp waits for another one.
template <bool WRITE_NPRED, bool DYNAMIC, int NC>
__global__ void __launch_bounds__(NC * 32, 1) fold_cash_mma_kernel(const FoldParams P, const MmaParams M) {
  // sources per pass over the k-steps and k-steps per loop iteration (= prefetch distance of the cube loads): ONE
  // variant of the pass, so that the hot code of all warps fits the 32 KB instruction cache; the register budget of
  // the 11-warp configuration only holds the loads of two k-steps
  constexpr int kPassSources = 2, kPassUnroll = NC <= 8 ? 4 : 2;
  extern __shared__ __align__(1024) unsigned char mma_smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  unsigned char* wp = mma_smem + (size_t)warp * kWarpBytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(mma_smem + M.lay_bars) + 2 * warp;  // this warp's two buffers
  unsigned char* tab_s = mma_smem + M.lay_tab;
  const double2* logtab = reinterpret_cast<const double2*>(mma_smem + M.lay_log);
  double* fs = reinterpret_cast<double*>(wp + kWpFs);

  // start-up: zero the exposure buffers (rows a dataset's boxes never write must read as finite numbers), tables
  for (int i = tid; i < NC * kWarpBytes / 16; i += NC * 32) reinterpret_cast<int4*>(mma_smem)[i] = make_int4(0, 0, 0, 0);
  if (M.tab_ds >= 0) {
    const int4* src = reinterpret_cast<const int4*>(P.ds[M.tab_ds].mma_table);
    for (int i = tid; i < M.tab_bytes / 16; i += NC * 32) reinterpret_cast<int4*>(tab_s)[i] = src[i];
  }
  for (int i = tid; i < kLogTabBytes / 16; i += NC * 32)
    reinterpret_cast<int4*>(mma_smem + M.lay_log)[i] = reinterpret_cast<const int4*>(M.logtab)[i];
  if (lane == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic zero fill before the async-proxy (TMA) writes
  __syncthreads();

  const int n_items = P.B * P.n_tiles;
  const int n_work = P.work ? *P.n_work : n_items;
  const int n_chunks = P.dl_bytes / 16;
  const int g = lane >> 2, tg = lane & 3;
  const uint64_t pol_keep = l2_policy_evict_last(), pol_stream = l2_policy_evict_first();
  const double log_trunc = P.log_trunc;
  const int n_warps_total = (int)gridDim.x * NC;

  auto claim = [&](int cur) -> int {  // a work position after `cur` (warp uniform)
    if (!DYNAMIC) return cur + n_warps_total;
    int nxt = 0;
    if (lane == 0) nxt = n_warps_total + (int)atomicAdd(P.counter, 1u);
    return __shfl_sync(0xffffffffu, nxt, 0);
  };
  auto load_desc = [&](int pos) -> int4 {  // one 16-byte chunk of the work descriptor per lane
    int4 v = make_int4(0, 0, 0, 0);
    if (pos < n_work && lane < n_chunks) v = __ldg(reinterpret_cast<const int4*>(P.desc + (size_t)pos * P.dl_bytes) + lane);
    return v;
  };
  // descriptor registers -> shared-memory slot, tile info computed by the whole warp (warp uniform values)
  auto install_tile = [&](int pos, const int4& dv, int slot) {
    if (lane < n_chunks) *reinterpret_cast<int4*>(wp + kWpDesc + slot * 448 + 16 * lane) = dv;
    const int item = P.work ? __ldg(P.work + pos) : pos;
    const int b = item % P.B, tile = item / P.B;
    const int d = __ldg(P.tile_ds + tile);
    const DatasetDev& D = P.ds[d];
    const int nx = D.nx, Pn = (int)D.P;
    const int nl = __shfl_sync(0xffffffffu, dv.x, 0), nover = __shfl_sync(0xffffffffu, dv.y, 0);
    const int evalidx = __shfl_sync(0xffffffffu, dv.z, 0);
    const int pt = (tile - D.tile_begin) * kFoldTPX;
    // which sub-items does a source reach?  lane = 4 sub + l tests source l (first 16 bytes of its InstDev: the
    // cutout rectangle) against sub-item sub
    const int l = lane & 3, sub = (lane >> 2) & 3;
    const int srcl = (P.dl_inst + l * (int)sizeof(InstDev)) / 16;
    const int rx = __shfl_sync(0xffffffffu, dv.x, srcl), ry = __shfl_sync(0xffffffffu, dv.y, srcl);
    const int rw = __shfl_sync(0xffffffffu, dv.z, srcl), rh = __shfl_sync(0xffffffffu, dv.w, srcl);
    const int p0 = pt + sub * kMmaSub, p1 = min(p0 + kMmaSub, Pn) - 1;
    const int y_first = p0 / nx, y_last = p1 / nx;
    const int x_first = p0 - y_first * nx, x_last = p1 - y_last * nx;
    bool ov = lane < 16 && l < nl && p0 < Pn && ry <= y_last && ry + rh > y_first;
    if (ov && y_first == y_last) ov = rx <= x_last && rx + rw > x_first;
    unsigned src_mask = __ballot_sync(0xffffffffu, ov);
    if (nover > 0) src_mask = 0xffffu;
    if (lane == 0) {
      const bool has_counts = D.counts != nullptr;
      MmaTileInfo ti;
      ti.w = pos;
      ti.tile = tile;
      ti.b = b;
      ti.d = d;
      ti.evalidx = evalidx;
      ti.nsrc = nl + nover;
      ti.src_mask = (int)src_mask;
      ti.pt = pt;
      ti.nx = nx;
      ti.nT = D.nT;
      ti.nR = D.nR;
      ti.Pn = Pn;
      ti.G = D.n_groups;
      ti.dflags = ((D.background != nullptr && P.include_background) ? kHdrHasBkg : 0) | (has_counts ? kHdrHasCounts : 0) |
                  ((has_counts && D.mask != nullptr) ? kHdrHasMask : 0);
      ti.pad0_ = ti.pad1_ = 0;
      *reinterpret_cast<MmaTileInfo*>(wp + kWpInfo + slot * 64) = ti;
    }
    __syncwarp();
  };
  // Start the loads of sub-item `sub` of the tile in `slot`: exposure into buffer `buf` (if a source reaches the
  // sub-item), epilogue inputs towards L2.  Returns whether the exposure is on its way.
  auto issue_sub = [&](int slot, int sub, int buf) -> bool {
    const MmaTileInfo& ti = *reinterpret_cast<const MmaTileInfo*>(wp + kWpInfo + slot * 64);
    const int p0 = ti.pt + sub * kMmaSub;
    const bool has_src = ((ti.src_mask >> (4 * sub)) & 15) != 0 && !(M.debug_skip & 4);
    __syncwarp();  // every lane is done reading the buffer's previous contents
    if (lane == 0) {
      const unsigned char* tm = reinterpret_cast<const unsigned char*>(P.ds[ti.d].tmaps_mma);
      if (!(M.debug_skip & 8)) {
        if (ti.dflags & kHdrHasBkg) tma_prefetch_2d(tm + 128, p0, 0);
        if (ti.dflags & kHdrHasCounts) tma_prefetch_2d(tm + 256, p0, 0);
        if (ti.dflags & kHdrHasMask) tma_prefetch_2d(tm + 384, p0, 0);
      }
      if (has_src) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive_expect_tx(&full[buf], (uint32_t)ti.nT * 128u);
        tma_load_2d(wp + kWpExp + buf * (kMmaMaxNT * 128), tm, p0, 0, &full[buf], pol_stream);
      }
    }
    return has_src;
  };

  // ---- prologue: this warp's first two tiles
  int w_cur = (int)blockIdx.x * NC + warp;
  if (w_cur >= n_work) return;
  int w_next = claim(w_cur);
  int slot = 0, buf = 0;
  uint32_t par0 = 0, par1 = 0;  // parity of the next completion of full[0] / full[1]
  install_tile(w_cur, load_desc(w_cur), 0);
  int4 dv_next = load_desc(w_next);
  bool cur_has_exp = issue_sub(0, 0, 0);

  for (;;) {
    // the tile after next is claimed now: its position is needed one tile later
    const int w_after = w_next < n_work ? claim(w_next) : w_next;
    const MmaTileInfo ti = *reinterpret_cast<const MmaTileInfo*>(wp + kWpInfo + slot * 64);
    const int nx = ti.nx, nT = ti.nT, nR = ti.nR, Pn = ti.Pn;
    bool more = true;
#pragma unroll 1
    for (int sub = 0; sub < kMmaSubs; ++sub) {
      const int p0 = ti.pt + sub * kMmaSub;
      if (p0 >= Pn) break;
      const int npx = min(kMmaSub, Pn - p0);
      // next sub-item (of this tile or of the next one): start its loads
      bool next_has_exp = false;
      if (sub + 1 < kMmaSubs && p0 + kMmaSub < Pn) {
        next_has_exp = issue_sub(slot, sub + 1, buf ^ 1);
      } else if (w_next < n_work) {
        install_tile(w_next, dv_next, slot ^ 1);
        next_has_exp = issue_sub(slot ^ 1, 0, buf ^ 1);
      } else {
        more = false;
      }
      const unsigned char* es = wp + kWpExp + buf * (kMmaMaxNT * 128);
      const bool qlive = 4 * g < npx;

      double acc[4][kMmaNTiles][2];
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int n = 0; n < kMmaNTiles; ++n) acc[j][n][0] = acc[j][n][1] = 0.0;

      if (cur_has_exp) {
        mbar_wait(&full[buf], buf ? par1 : par0);
        if (buf) par1 ^= 1u; else par0 ^= 1u;
      }
      if (cur_has_exp && !(M.debug_skip & 1)) {
        const unsigned char* dsc = wp + kWpDesc + slot * 448;
        const int ntot = ti.nsrc;
        const InstDev* slist = reinterpret_cast<const InstDev*>(dsc + P.dl_inst);
        const int* over = reinterpret_cast<const int*>(dsc + P.dl_over);
        const unsigned char* tab = ti.d == M.tab_ds ? tab_s : reinterpret_cast<const unsigned char*>(P.ds[ti.d].mma_table);
        const MmaTableHead* th = reinterpret_cast<const MmaTableHead*>(tab);
        const double* frag_all = reinterpret_cast<const double*>(tab + sizeof(MmaTableHead));
        const int pq = p0 + 4 * g;
        const int qy = pq / nx, qx = pq - qy * nx;
        const float* zf = P.zeros;
        for (int gi = 0; gi < ti.G; ++gi) {
          const unsigned long long bits = th->bits[gi];
          const int ks_lo = th->ks_lo[gi], ks_hi = th->ks_hi[gi];
          const double* tabg = frag_all + (size_t)gi * kMmaGroupDoubles;
          int idx = 0;
          while (idx < ntot) {
            MmaSrc src;
            int foff[kPassSources];
            int na = 0;
#pragma unroll
            for (int l = 0; l < kPassSources; ++l) {
              src.cptr[l] = zf;
              src.pstr[l] = 0;
              foff[l] = -1;
              const InstDev* in = nullptr;
              while (idx < ntot) {  // next source of this edisp group (warp uniform)
                const InstDev* cand = idx < kFoldLMax ? slist + idx : P.insts + over[idx - kFoldLMax];
                ++idx;
                if (ti.G == 1 || cand->group == gi) {
                  in = cand;
                  break;
                }
              }
              if (in) {
                ++na;
                foff[l] = in->flux_off;
                const int ly = qy - in->y0, lx = qx - in->x0a;
                if (qlive && ly >= 0 && ly < in->cy && lx >= 0 && lx < in->pitch) {
                  src.cptr[l] = in->conv + ly * in->pitch + lx;
                  src.pstr[l] = in->plane_stride;
                }
              }
            }
            if (na == 0) break;
            // this pass' flux rows, zero padded to 64 true bins, into the warp's scratch
            __syncwarp();
#pragma unroll
            for (int l = 0; l < kPassSources; ++l) {  // (rows without a source multiply zeros: any finite value will do)
              fs[l * kMmaMaxNT + lane] = l < na && lane < nT ? __ldg(P.flux + foff[l] + lane) : 0.0;
              fs[l * kMmaMaxNT + 32 + lane] = l < na && lane + 32 < nT ? __ldg(P.flux + foff[l] + 32 + lane) : 0.0;
            }
            __syncwarp();
            fold_pass<kPassSources, kPassUnroll>(acc, src, fs, es, tabg, bits, ks_lo & ~(kPassUnroll - 1), ks_hi, nT, g, tg,
                                                 lane, pol_keep);
          }
        }
      }

      // ---- epilogue: background model, clip, Cash.  Thread: pixels p0 + 4g + j, reco bins 8n + 2tg + i.  Its inputs
      // come straight from global memory (this warp's TMA prefetch put them in L2 one sub-item ago), two n-tiles
      // (four rows) a round.
      if (!(M.debug_skip & 2)) {
        EpiArgs ea;
        ea.bkgp = P.ds[ti.d].background;
        ea.cntp = P.ds[ti.d].counts;
        ea.mskp = P.ds[ti.d].mask;
        ea.bf = P.bkgfac + (size_t)ti.evalidx * P.smem_nRp;
        ea.npred_out = P.npred_out;
        ea.logtab = logtab;
        ea.log_trunc = log_trunc;
        ea.pol = pol_stream;
        ea.Pn = Pn;
        ea.col = qlive ? p0 + 4 * g : p0;
        ea.nR = nR;
        ea.live = qlive;
        ea.tg = tg;
        ea.flags = ti.dflags;
        ea.include_background = P.include_background;
        double stat = mma_epilogue<WRITE_NPRED, true>(acc, ea);
        if (!WRITE_NPRED) {
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) stat += __shfl_xor_sync(0xffffffffu, stat, o);
          if (!(fabs(stat) <= 1.7976931348623157e308)) {  // warp uniform; inf / nan inputs: the reference's log decides
            stat = mma_epilogue<WRITE_NPRED, false>(acc, ea);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) stat += __shfl_xor_sync(0xffffffffu, stat, o);
          }
          if (lane == 0) P.partials[((size_t)ti.b * P.n_tiles + ti.tile) * kMmaSubs + sub] = stat;
        }
      }
      cur_has_exp = next_has_exp;
      buf ^= 1;
    }
    if (!more) break;
    slot ^= 1;
    w_cur = w_next;
    w_next = w_after;
    dv_next = load_desc(w_next);
  }
}

}  // namespace gpy
