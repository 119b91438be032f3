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
// consumer warps per CTA (+ 1 producer warp): 11 -> 384 threads x 168 registers, 7 -> 256 threads x 255 registers
// (ptxas sizes the register budget for thread counts rounded up to 128)

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
  int n_stages, stage_bytes;     // ring of tile stages: [nT][512 B] exposure tile + 448 B descriptor + 64 B tile info
  unsigned lay_tab, lay_log, lay_fs, lay_bars;  // shared-memory offsets behind the ring: cached table, log table,
                                                // per-warp flux rows, mbarriers (+ sequence counter)
  int tab_bytes;                 // bytes of the table cached in shared memory (dataset tab_ds)
  int tab_ds;                    // dataset whose table is cached (-1: none)
  const double2* logtab;         // fast_log table (device)
  int debug_skip;                // GPY_MMA_SKIP (profiling only): 1 skips the k-loops, 2 the epilogue, 4 the exposure loads
  long long* trace;              // GPY_MMA_TRACE (profiling only): [0] = event count, then {type, warp, seq, clock} of CTA 0
};
#define GPY_TRACE(type, seq)                                                                      \
  do {                                                                                            \
    if (M.trace && blockIdx.x == 0 && lane == 0) {                                                \
      const unsigned long long i_ = atomicAdd(reinterpret_cast<unsigned long long*>(M.trace), 1ull); \
      if (i_ < 60000) {                                                                           \
        M.trace[1 + 4 * i_] = (type);                                                             \
        M.trace[2 + 4 * i_] = warp;                                                               \
        M.trace[3 + 4 * i_] = (seq);                                                              \
        M.trace[4 + 4 * i_] = clock64();                                                          \
      }                                                                                           \
    }                                                                                             \
  } while (0)

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
// iteration (UN k-steps): the zero-padded flux rows fs[l][64] and the dense fragment table tabg[ks][n][32] in shared
// memory, the cube rows t, t+1, t+8, t+9 (UN = 4) through 32-bit byte steps; the exposure tile [nT][128 px] is read
// at this thread's quad (the four rows of a k-step share their banks: a 4-way conflict on one LDS.128 per k-step).  The cube loads of the next iteration are issued before the DMMAs of this one (PD = UN k-steps ahead).
template <int NL, int UN>
__device__ __forceinline__ void fold_pass(double (&acc)[4][kMmaNTiles][2], const MmaSrc& S, const double* fs,
                                          const unsigned char* es, const double* tabg, unsigned long long bits,
                                          int k_begin, int k_end, int nT, int g, int tg, int lane, uint64_t pol) {
  static_assert(UN == 2 || UN == 4, "k-steps per iteration");
  const int t_lane = 2 * tg;
  // per-lane constants (es: this thread's pixel quad in row 0 of the exposure tile, rows of kFoldTPX floats)
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
      // (rows past the axis read the last row: finite, and their flux is zero)
      const float4 e = *reinterpret_cast<const float4*>(es + min(t0 + blk * 8 + h, nT - 1) * (kFoldTPX * 4));
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

// Per-tile information in a stage of the ring.
struct MmaTileInfo {
  int w, tile, b, d;            // work position (descriptor slot; -1: no work left), tile, parameter vector, dataset
  int evalidx, nsrc, src_mask, pt;  // (dataset, vector) index; overlapping sources; bit 4 sub + l: source l reaches
                                    // sub-item sub; first pixel of the tile in its dataset
  int nx, nT, nR, Pn;
  int G, dflags, pad0_, pad1_;
};
static_assert(sizeof(MmaTileInfo) == 64, "MmaTileInfo");

// One CTA per SM: NC consumer warps + 1 producer warp, no CTA barrier after start-up.
//   * The producer claims tiles (128 pixels x one parameter vector) from a global counter two tiles ahead, reads their
//     work descriptors one tile ahead, and fills the stages of a shared-memory ring: the exposure tile [nT][128 px]
//     by ONE 2-D tensor-map TMA load (512-byte rows: fewer, longer DRAM bursts and a quarter of the TMA instructions
//     of per-sub-item boxes), the descriptor and the tile info; the background / counts / mask boxes of the tile are
//     sent towards L2 with TMA prefetches at the same time.
//   * Consumer warps take (tile, sub-item) pairs in sequence order from a shared counter; a stage is free again when
//     the four sub-items of its tile are past their k-loops (empty barrier, count 4).
template <bool WRITE_NPRED, bool DYNAMIC, int NC>
__global__ void __launch_bounds__((NC + 1) * 32, 1) fold_cash_mma_kernel(const FoldParams P, const MmaParams M) {
  // sources per pass over the k-steps and k-steps per loop iteration (= prefetch distance of the cube loads): ONE
  // variant of the pass, so that the hot code of all warps fits the 32 KB instruction cache; the register budget of
  // the 11-warp configuration only holds the loads of two k-steps
  constexpr int kPassSources = 2, kPassUnroll = NC <= 7 ? 4 : 2;
  constexpr int kThreads = (NC + 1) * 32;
  extern __shared__ __align__(1024) unsigned char mma_smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int S = M.n_stages, stage_bytes = M.stage_bytes, exp_bytes = M.stage_bytes - 512;
  uint64_t* full = reinterpret_cast<uint64_t*>(mma_smem + M.lay_bars);
  uint64_t* empty = full + S;
  int* seq_next = reinterpret_cast<int*>(empty + S);
  unsigned char* tab_s = mma_smem + M.lay_tab;
  const double2* logtab = reinterpret_cast<const double2*>(mma_smem + M.lay_log);
  double* fs = reinterpret_cast<double*>(mma_smem + M.lay_fs) + warp * (kPassSources * kMmaMaxNT);

  // start-up: tables, barriers
  if (M.tab_ds >= 0) {
    const int4* src = reinterpret_cast<const int4*>(P.ds[M.tab_ds].mma_table);
    for (int i = tid; i < M.tab_bytes / 16; i += kThreads) reinterpret_cast<int4*>(tab_s)[i] = src[i];
  }
  for (int i = tid; i < kLogTabBytes / 16; i += kThreads)
    reinterpret_cast<int4*>(mma_smem + M.lay_log)[i] = reinterpret_cast<const int4*>(M.logtab)[i];
  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kMmaSubs);
    }
    *seq_next = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int n_items = P.B * P.n_tiles;
  const int n_work = P.work ? *P.n_work : n_items;
  const int n_chunks = P.dl_bytes / 16;

  if (warp == NC) {
    // ------------------------------------------------------------------------------------------------ producer
    const uint64_t pol_stream = l2_policy_evict_first();
    auto claim = [&](int cur) -> int {  // a work position after `cur` (warp uniform)
      if (!DYNAMIC) return cur + (int)gridDim.x;
      int nxt = 0;
      if (lane == 0) nxt = (int)gridDim.x + (int)atomicAdd(P.counter, 1u);
      return __shfl_sync(0xffffffffu, nxt, 0);
    };
    auto load_desc = [&](int pos) -> int4 {  // one 16-byte chunk of the work descriptor per lane
      int4 v = make_int4(0, 0, 0, 0);
      if (pos < n_work && lane < n_chunks) v = __ldg(reinterpret_cast<const int4*>(P.desc + (size_t)pos * P.dl_bytes) + lane);
      return v;
    };
    int stage = 0;
    uint32_t empty_par = 1;
    bool first_lap = true;
    auto next_stage = [&]() {
      if (++stage == S) {
        stage = 0;
        empty_par ^= 1u;
        first_lap = false;
      }
    };
    // Everything with a long latency is requested one iteration (tile) before it is used: the claim of the tile after
    // next (the atomic's result stays in lane 0 until the next iteration broadcasts it), the next tile's descriptor,
    // item and dataset index.  The dataset's fields are re-read only when the dataset changes.
    auto claim_async = [&]() -> int {  // lane 0: the claimed position; other lanes: garbage until claim_get
      int nxt = 0;
      if (DYNAMIC && lane == 0) nxt = (int)gridDim.x + (int)atomicAdd(P.counter, 1u);
      return nxt;
    };
    auto claim_get = [&](int pending, int cur) -> int {
      return DYNAMIC ? __shfl_sync(0xffffffffu, pending, 0) : cur + (int)gridDim.x;
    };
    auto load_item = [&](int pos) -> int { return pos < n_work ? (P.work ? __ldg(P.work + pos) : pos) : 0; };
    int w_cur = (int)blockIdx.x;
    int w_next = w_cur < n_work ? claim(w_cur) : w_cur;
    int pending = claim_async();  // the tile after next
    int4 dv = load_desc(w_cur);
    int item = load_item(w_cur);
    int d = w_cur < n_work ? __ldg(P.tile_ds + item / P.B) : 0;
    int item_next = load_item(w_next);
    int cached_d = -1;
    int nx = 1, Pn = 0, nT = 0, nR = 0, tile_begin = 0, G = 1, dflags = 0;
    const unsigned char* tm = nullptr;
    while (w_cur < n_work) {
      const int4 dv_next = load_desc(w_next);
      const int d_next = w_next < n_work ? __ldg(P.tile_ds + item_next / P.B) : 0;
      const int b = item % P.B, tile = item / P.B;
      if (d != cached_d) {
        const DatasetDev& D = P.ds[d];
        nx = D.nx;
        Pn = (int)D.P;
        nT = D.nT;
        nR = D.nR;
        tile_begin = D.tile_begin;
        G = D.n_groups;
        const bool has_counts = D.counts != nullptr;
        dflags = ((D.background != nullptr && P.include_background) ? kHdrHasBkg : 0) | (has_counts ? kHdrHasCounts : 0) |
                 ((has_counts && D.mask != nullptr) ? kHdrHasMask : 0);
        tm = reinterpret_cast<const unsigned char*>(D.tmaps);
        cached_d = d;
      }
      const int nl = __shfl_sync(0xffffffffu, dv.x, 0), nover = __shfl_sync(0xffffffffu, dv.y, 0);
      const int evalidx = __shfl_sync(0xffffffffu, dv.z, 0);
      const int pt = (tile - tile_begin) * kFoldTPX;
      // which sub-items does a source reach?  lane = 4 sub + l tests source l (first 16 bytes of its InstDev: the
      // cutout rectangle) against sub-item sub
      unsigned src_mask;
      {
        const int l = lane & 3, sub = (lane >> 2) & 3;
        const int srcl = (P.dl_inst + l * (int)sizeof(InstDev)) / 16;
        const int rx = __shfl_sync(0xffffffffu, dv.x, srcl), ry = __shfl_sync(0xffffffffu, dv.y, srcl);
        const int rw = __shfl_sync(0xffffffffu, dv.z, srcl), rh = __shfl_sync(0xffffffffu, dv.w, srcl);
        const int p0 = pt + sub * kMmaSub, p1 = min(p0 + kMmaSub, Pn) - 1;
        const int y_first = p0 / nx, y_last = p1 / nx;
        const int x_first = p0 - y_first * nx, x_last = p1 - y_last * nx;
        bool ov = lane < 16 && l < nl && p0 < Pn && ry <= y_last && ry + rh > y_first;
        if (ov && y_first == y_last) ov = rx <= x_last && rx + rw > x_first;
        src_mask = __ballot_sync(0xffffffffu, ov);
        if (nover > 0) src_mask = 0xffffu;
        if (M.debug_skip & 4) src_mask = 0u;
      }
      unsigned char* st = mma_smem + (size_t)stage * stage_bytes;
      GPY_TRACE(1, w_cur);
      if (!first_lap) mbar_wait(&empty[stage], empty_par);
      GPY_TRACE(2, w_cur);
      if (lane < n_chunks) *reinterpret_cast<int4*>(st + exp_bytes + 16 * lane) = dv;
      if (lane == 28) *reinterpret_cast<int4*>(st + exp_bytes + 448) = make_int4(w_cur, tile, b, d);
      if (lane == 29) *reinterpret_cast<int4*>(st + exp_bytes + 464) = make_int4(evalidx, nl + nover, (int)src_mask, pt);
      if (lane == 30) *reinterpret_cast<int4*>(st + exp_bytes + 480) = make_int4(nx, nT, nR, Pn);
      if (lane == 31) *reinterpret_cast<int4*>(st + exp_bytes + 496) = make_int4(G, dflags, 0, 0);
      __syncwarp();
      if (lane == 0) {
        if (!(M.debug_skip & 8)) {  // epilogue inputs towards L2; the consumers read them straight from global memory
          if (dflags & kHdrHasBkg) tma_prefetch_2d(tm + 128, pt, 0);
          if (dflags & kHdrHasCounts) tma_prefetch_2d(tm + 256, pt, 0);
          if (dflags & kHdrHasMask) tma_prefetch_2d(tm + 384, pt, 0);
        }
        if (src_mask) {
          mbar_arrive_expect_tx(&full[stage], (uint32_t)nT * (kFoldTPX * 4));
          tma_load_2d(st, tm, pt, 0, &full[stage], pol_stream);
        } else {
          mbar_arrive(&full[stage]);
        }
      }
      GPY_TRACE(3, w_cur);
      next_stage();
      w_cur = w_next;
      w_next = claim_get(pending, w_next);
      pending = claim_async();
      dv = dv_next;
      item = item_next;
      d = d_next;
      item_next = load_item(w_next);
    }
    // sentinel tiles: one sequence number per consumer warp
    for (int k = 0; k < (NC + kMmaSubs - 1) / kMmaSubs; ++k) {
      unsigned char* st = mma_smem + (size_t)stage * stage_bytes;
      if (!first_lap) mbar_wait(&empty[stage], empty_par);
      if (lane == 0) {
        *reinterpret_cast<int4*>(st + exp_bytes + 448) = make_int4(-1, 0, 0, 0);
        mbar_arrive(&full[stage]);
      }
      __syncwarp();
      next_stage();
    }
    return;
  }

  // -------------------------------------------------------------------------------------------------- consumers
  const int g = lane >> 2, tg = lane & 3;
  const uint64_t pol_keep = l2_policy_evict_last(), pol_stream = l2_policy_evict_first();
  const double log_trunc = P.log_trunc;
  for (;;) {
    int seq = 0;
    if (lane == 0) seq = atomicAdd(seq_next, 1);
    seq = __shfl_sync(0xffffffffu, seq, 0);
    const int tseq = seq >> 2, sub = seq & 3;
    const int s = tseq % S;
    unsigned char* st = mma_smem + (size_t)s * stage_bytes;
    GPY_TRACE(10, seq);
    mbar_wait(&full[s], (uint32_t)(tseq / S) & 1u);
    GPY_TRACE(11, seq);
    const MmaTileInfo ti = *reinterpret_cast<const MmaTileInfo*>(st + exp_bytes + 448);
    if (ti.w < 0) break;  // no work left (nobody reuses a sentinel stage)
    const int nx = ti.nx, nT = ti.nT, nR = ti.nR, Pn = ti.Pn;
    const int p0 = ti.pt + sub * kMmaSub;
    if (p0 >= Pn) {  // the dataset ends inside this tile
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
      continue;
    }
    const int npx = min(kMmaSub, Pn - p0);
    const bool qlive = 4 * g < npx;

    double acc[4][kMmaNTiles][2];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int n = 0; n < kMmaNTiles; ++n) acc[j][n][0] = acc[j][n][1] = 0.0;

    if (((ti.src_mask >> (4 * sub)) & 15) != 0 && !(M.debug_skip & 1)) {
      const unsigned char* es = st + sub * (kMmaSub * 4) + g * 16;  // this thread's quad in row 0 of the exposure tile
      const unsigned char* dsc = st + exp_bytes;
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
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);  // this sub-item is done with the stage (exposure tile, descriptor, info)
    GPY_TRACE(12, seq);

    // ---- epilogue: background model, clip, Cash.  Thread: pixels p0 + 4g + j, reco bins 8n + 2tg + i.  Its inputs
    // come straight from global memory (the producer's TMA prefetch put them in L2), two n-tiles (four rows) a round.
    if (M.debug_skip & 2) continue;
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
    GPY_TRACE(13, seq);
  }
}

}  // namespace gpy
