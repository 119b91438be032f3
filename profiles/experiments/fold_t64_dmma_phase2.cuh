// fold_cash_t64_kernel: the fused fold + Cash kernel on 64-pixel sub-tiles, three CTAs per SM.
//
// Same work items, descriptors, tensor-map staging protocol and arithmetic as fold_cash_tma_kernel (kernels.cuh); what
// changes is the shape of a CTA's working set.  fold_cash_tma_kernel holds a 128-pixel tile (110 KB of shared memory,
// 128 registers: two CTAs per SM, and one-versus-two CTAs per SM measured 373 vs 233 us at C3: the kernel lives on the
// other CTA filling its stalls).  Here a CTA works on HALF a tile at a time:
//   * all true-energy bins (<= 64) in ONE phase-1 pass: thread = pixel quad (16) x energy slot (16), 4 rows each,
//     v[t][64 px] in fp64 = 30 KB at nT = 60;
//   * phase 2: thread = pixel (64) x reco chunk (4): 8 accumulators instead of 16 -- phase 1 and phase 2 no longer
//     overlap in registers either (no accumulators live across passes);
//   * tensor copies of [rows][64] boxes: 15 KB exposure + 17 KB epilogue inputs;
// 75 KB of shared memory and <= 80 registers: three CTAs (24 warps) per SM, three independent barrier domains.
// The two halves of an item share its descriptor, flux rows and partial sum.
// Measured against fold_cash_tma_kernel (fold kernel only): C3 227 vs 232 us, c3-flat 238 vs 245, one C5-shaped dataset
// 657 vs 695, the 64-observation joint evaluation 2553 vs 2713, C1 16.2 vs 18.3.  Two variants were tried and are not the
// default: double-buffered inputs at two CTAs per SM (STAGES = 2: 264 us -- the copies' latency is not what the
// sub-tiles wait for, occupancy is) and a fused epilogue that takes the Cash terms straight from the accumulators
// (no parking, two barriers less per sub-tile: 321 us -- the warps of the low-energy chunks take all the logs).
// Eligible launches (host): tensor maps on every dataset, one edisp group, nT <= 64, nR <= 32.
#pragma once
#include "kernels.cuh"

namespace gpy {

#ifndef GPY_T64_DMMA  // phase 2 on the fp64 tensor-core instruction (mma.sync.m8n8k4.f64: the DFMA pipe's rate, a third of the issue slots)
#define GPY_T64_DMMA 1
#endif
#ifndef GPY_T64_P2_PAIR  // phase 2: thread = two pixels x four reco bins (else one pixel x eight)
#define GPY_T64_P2_PAIR 1
#endif
constexpr int kT64 = 64;          // pixels per sub-tile
constexpr int kT64Slots = 16;     // true-energy slots of phase 1 (256 threads / 16 quads)
constexpr int kT64Rows = 4;       // rows per thread: t = slot + 16 k
constexpr int kT64Log = 64;       // fast_log table entries kept in shared memory (|r| < 2^-7: degree-7 log1p still < 1e-17)

// byte strides between the two stages of a double-buffered input (also the sizes of one stage)
__host__ __device__ inline size_t t64_stride_e(int nT) { return (size_t)nT * kT64 * 4; }
__host__ __device__ inline size_t t64_stride_b(int nR) { return (size_t)nR * kT64 * 4; }
__host__ __device__ inline size_t t64_stride_c(int nR, bool compact) { return (size_t)nR * kT64 * (compact ? 2 : 4); }
__host__ __device__ inline size_t t64_stride_m(int nR, bool compact) {
  return compact ? ((size_t)nR * 16 + 127) / 128 * 128 : ((size_t)nR * kT64 + 127) / 128 * 128;  // tensor-copy targets: 128-byte aligned
}

__host__ __device__ inline FoldSmem fold_t64_layout(int nT, int nR, int nRp, int pdfc_rows, bool compact, int stages) {
  FoldSmem L;
  size_t o = 0;
  L.es = o;    o += stages * t64_stride_e(nT);
  L.bs = o;    o += stages * t64_stride_b(nR);
  L.cs = o;    o += stages * t64_stride_c(nR, compact);
  L.ms = o;    o += stages * t64_stride_m(nR, compact);
  const size_t vrows = (size_t)(nT > nRp ? nT : nRp);  // v[t][64] doubles; mu[r][64] parked in the same buffer
  L.vs = o;    o += vrows * kT64 * 8;
  L.pdf = o;   o += (size_t)pdfc_rows * 64;
  L.desc = o;  o += (size_t)2 * desc_layout().bytes;
  L.dyn = o;   o += (size_t)2 * (kFoldLMax * nT + nRp) * 8;
  L.bars = (o + 7) / 8 * 8;
  L.total = L.bars + 6 * 8;  // exposure x 2 stages, epilogue inputs x 2 stages, descriptor x 2 buffers
  return L;
}

// D (8 x 8) += A (8 x 4, row major) * B (4 x 8, column major) in fp64: a = A[lane / 4][lane % 4], b = B[lane % 4][lane / 4],
// d0 / d1 = D[lane / 4][2 (lane % 4) + 0 / 1].
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// log(x) with the 6-bit table (see fast_log in kernels.cuh)
__device__ __forceinline__ double fast_log64(double x, const double2* __restrict__ tab) {
  const long long ix = __double_as_longlong(x);
  const int e = (int)(ix >> 52) - 1023;
  const double2 t = tab[(int)(ix >> 46) & 63];
  const double m = __longlong_as_double((ix & 0x000fffffffffffffLL) | 0x3ff0000000000000LL);
  const double r = fma(m, t.x, -1.0);
  const double r2 = r * r;
  const double a = fma(r, 1.0 / 3.0, -0.5), b = fma(r, 0.2, -0.25), c = fma(r, 1.0 / 7.0, -1.0 / 6.0);
  double q = fma(r2, c, b);
  q = fma(r2, q, a);
  const double p = fma(r2, q, r);
  return fma((double)e, 0.6931471805599453, t.y) + p;
}

// STAGES = 1: single-buffered inputs, three CTAs per SM (the next exposure is requested after phase 1, the next
// epilogue inputs after the end-of-sub-tile barrier).  STAGES = 2: double-buffered inputs, two CTAs per SM: the copies
// of sub-tile k + 1 go out at the top of sub-tile k and have its whole duration to land.
template <bool WRITE_NPRED, bool DYNAMIC, bool COMPACT, int STAGES>
__global__ void __launch_bounds__(kFoldThreads, STAGES == 1 ? 3 : 2) fold_cash_t64_kernel(const FoldParams P) {
  extern __shared__ __align__(128) unsigned char fold_raw[];
  constexpr bool DB = STAGES == 2;
  unsigned char* es_base = fold_raw + P.lay_es;
  unsigned char* bs_base = fold_raw + P.lay_bs;
  unsigned char* cs_base = fold_raw + P.lay_cs;
  unsigned char* ms_base = fold_raw + P.lay_ms;
  const uint32_t st_e = (uint32_t)t64_stride_e(P.smem_nT), st_b = (uint32_t)t64_stride_b(P.smem_nR),
                 st_c = (uint32_t)t64_stride_c(P.smem_nR, COMPACT), st_m = (uint32_t)t64_stride_m(P.smem_nR, COMPACT);
  double* vs = reinterpret_cast<double*>(fold_raw + P.lay_vs);
  double* pdf_s = reinterpret_cast<double*>(fold_raw + P.lay_pdf);
  unsigned char* desc_s = fold_raw + P.lay_desc;
  double* dyn_s = reinterpret_cast<double*>(fold_raw + P.lay_dyn);
  DescLayout DL;
  DL.inst = P.dl_inst;
  DL.over = P.dl_over;
  DL.bytes = P.dl_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(fold_raw + P.lay_bars);  // [0, 1] exposure stages, [2, 3] epilogue-input stages, [4, 5] descriptors
  __shared__ double wsum[kFoldThreads / 32];
  __shared__ int2 nxt_s;
  __shared__ DatasetDev Ds;
  __shared__ __align__(16) int band_s[16];
  __shared__ double2 logtab_s[kT64Log];
  if (threadIdx.x < kT64Log) logtab_s[threadIdx.x] = P.logtab64[threadIdx.x];

  const int tid = threadIdx.x;
  const int n_items = P.B * P.n_tiles;
  const uint64_t pol_keep = l2_policy_evict_last(), pol_stream = l2_policy_evict_first();
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&bars[2], 1);
    mbar_init(&bars[3], 1);
    mbar_init(&bars[4], 1);
    mbar_init(&bars[5], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  constexpr int kEpiTid = 32, kClaimTid = 64, kExpoTid = 96;
  auto geom = [&](int item, int& p0, int& npx) -> const DatasetDev* {
    const int tile = DYNAMIC ? item : item / P.B;
    const DatasetDev* Dn = &Ds;
    if (tile < Ds.tile_begin || tile >= Ds.tile_begin + Ds.n_tiles) Dn = P.ds + P.tile_ds[tile];
    p0 = (tile - Dn->tile_begin) * kFoldTPX;
    npx = min(kFoldTPX, (int)Dn->P - p0);
    return Dn;
  };
  // tensor copies of sub-tile h of `item` (64-pixel boxes: the second set of four tensor maps of a dataset)
  auto issue_expo = [&](int item, int h, uint32_t st) {
    if (tid != kExpoTid) return;
    int p0, npx;
    const DatasetDev* D = geom(item, p0, npx);
    mbar_arrive_expect_tx(&bars[st], (uint32_t)kT64 * 4u * (uint32_t)D->nT);
    tma_load_2d(es_base + st * st_e, reinterpret_cast<const unsigned char*>(D->tmaps) + 512, p0 + kT64 * h, 0, &bars[st],
                pol_stream);
  };
  auto issue_epi = [&](int item, int h, uint32_t st) {
    if (tid != kEpiTid) return;
    float* bs = reinterpret_cast<float*>(bs_base + st * st_b);
    float* cs = reinterpret_cast<float*>(cs_base + st * st_c);
    unsigned char* ms = ms_base + st * st_m;
    uint64_t* bar = &bars[2 + st];
    int p0, npx;
    const DatasetDev* D = geom(item, p0, npx);
    const unsigned char* tm = reinterpret_cast<const unsigned char*>(D->tmaps) + 512;
    const bool has_bkg = D->background && P.include_background, has_counts = D->counts != nullptr,
               has_mask = has_counts && D->mask;
    const uint32_t nR = (uint32_t)D->nR;
    uint32_t total = 0;
    if (has_bkg) total += (uint32_t)kT64 * 4u * nR;
    if (has_counts) total += (uint32_t)kT64 * (COMPACT ? 2u : 4u) * nR;
    if (has_mask) total += (COMPACT ? 16u : (uint32_t)kT64) * nR;
    mbar_arrive_expect_tx(bar, total);
    const int ph = p0 + kT64 * h;
    if (has_bkg) tma_load_2d(bs, tm + 128, ph, 0, bar, pol_stream);
    if (has_counts) tma_load_2d(cs, tm + 256, ph, 0, bar, pol_stream);
    // (compact: the 16 bytes of mask bits of the whole 128-pixel tile -- a tensor-copy row cannot be shorter)
    if (has_mask) tma_load_2d(ms, tm + 384, COMPACT ? p0 >> 3 : ph, 0, bar, pol_stream);
  };
  auto fetch_desc = [&](int it, uint32_t buf) {
    if (tid != kClaimTid) return;
    mbar_arrive_expect_tx(&bars[4 + buf], (uint32_t)DL.bytes);
    bulk_g2s(desc_s + buf * DL.bytes, P.desc + (size_t)it * DL.bytes, (uint32_t)DL.bytes, &bars[4 + buf]);
  };
  const int dyn_len = kFoldLMax * P.smem_nT + P.smem_nRp;
  auto load_dyn = [&](uint32_t buf, int i) -> double {
    const unsigned char* dsc = desc_s + buf * DL.bytes;
    const int4 head = *reinterpret_cast<const int4*>(dsc);
    const InstDev* inst = reinterpret_cast<const InstDev*>(dsc + DL.inst);
    const int l = (i >= P.smem_nT) + (i >= 2 * P.smem_nT) + (i >= 3 * P.smem_nT) + (i >= 4 * P.smem_nT);
    const int t = i - l * P.smem_nT;
    if (l == kFoldLMax) return __ldg(P.bkgfac + (size_t)head.z * P.smem_nRp + t);
    if (l < head.x) return __ldg(P.flux + inst[l].flux_off + t);
    return 0.0;
  };
  const int n_work = P.work ? *P.n_work : n_items;
  auto item_at = [&](int pos) -> int { return pos < n_work ? (P.work ? __ldg(P.work + pos) : pos) : n_items; };
  const int* zero_i = reinterpret_cast<const int*>(P.zeros);
  auto load_dataset = [&](int tile) {
    const int* src = reinterpret_cast<const int*>(P.ds + P.tile_ds[tile]);
    if (tid < (int)(sizeof(DatasetDev) / 4)) reinterpret_cast<int*>(&Ds)[tid] = src[tid];
    __syncthreads();
    const double* pdfc = Ds.pdfc;
    for (int i = tid; i < Ds.pdfc_rows * 8; i += kFoldThreads) pdf_s[i] = pdfc[i];
    const int* band = Ds.band;
    if (tid < 16) band_s[tid] = tid < (Ds.nRp >> 3) * 4 ? band[tid] : 0;
  };

  int w = blockIdx.x;
  int item = item_at(w);
  if (item >= n_items) return;
  int wn, next;
  if (!DYNAMIC) {
    wn = w + gridDim.x;
    next = item_at(wn);
  } else if (tid == kClaimTid) {
    const int c = gridDim.x + atomicAdd(P.counter, 1u);
    nxt_s = make_int2(c, item_at(c));
  }
  load_dataset(DYNAMIC ? item : item / P.B);
  __syncthreads();
  if (DYNAMIC) {
    const int2 v = nxt_s;
    wn = v.x;
    next = v.y;
  }
  {
    if (__ldg(reinterpret_cast<const int*>(P.desc + (size_t)w * DL.bytes)) > 0) issue_expo(item, 0, 0);
    issue_epi(item, 0, 0);
    fetch_desc(w, 0);
    mbar_wait(&bars[4], 0);
    for (int i = tid; i < dyn_len; i += kFoldThreads) dyn_s[i] = load_dyn(0, i);
    __syncthreads();
  }
  uint32_t it = 0;        // items done: descriptor buffer it & 1, its mbarrier phase (it >> 1) & 1
  uint32_t eparity = 0;   // exposure barriers: bit s = phase of stage s, flips per sub-tile with sources that used it
  uint32_t sparity = 0;   // epilogue-input barriers: bit s = phase of stage s, flips per sub-tile that used it
  uint32_t ksub = 0;      // sub-tiles done by this CTA: stage ksub & 1 when double buffered

  const int q = tid & 15, slot = tid >> 4;    // phase 1: 16 pixel quads x 16 energy slots
#if GPY_T64_P2_PAIR
  const int px = (tid & 31) * 2, cslot = tid >> 6, rh = ((tid >> 5) & 1) * 4;  // phase 2: 32 pixel pairs x 4 reco chunks x 2 halves of a chunk
#else
  const int px = tid & 63, cslot = tid >> 6;  // phase 2: 64 pixels x 4 reco chunks
#endif

  for (;; ++it) {
    const uint32_t parity = it & 1u;
    const uint32_t dphase = (it >> 1) & 1u;
    const int b = DYNAMIC ? 0 : item % P.B;
    const int tile = DYNAMIC ? item : item / P.B;
    int wnn = 0, nn = n_items;
    unsigned claim_raw = 0;
    if (!DYNAMIC) {
      wnn = wn + gridDim.x;
      nn = item_at(wnn);
    } else if (tid == kClaimTid) {
      claim_raw = atomicAdd(P.counter, 1u);
    }
    const int nl_next = __ldg(next < n_items ? reinterpret_cast<const int*>(P.desc + (size_t)wn * DL.bytes) : zero_i);
    if (next < n_items) fetch_desc(wn, parity ^ 1u);
    if (tile < Ds.tile_begin || tile >= Ds.tile_begin + Ds.n_tiles) {
      __syncthreads();
      load_dataset(tile);
      __syncthreads();
    }
    const int nT = Ds.nT, nR = Ds.nR, nx = Ds.nx;
    const int p0 = (tile - Ds.tile_begin) * kFoldTPX;
    const int npx = min(kFoldTPX, (int)Ds.P - p0);
    const bool has_bkg = Ds.background && P.include_background, has_counts = Ds.counts != nullptr,
               has_mask = Ds.mask != nullptr;
    const int nchunks = Ds.nRp >> 3;
    mbar_wait(&bars[4 + parity], dphase);
    const unsigned char* dsc = desc_s + parity * DL.bytes;
    const int nl = reinterpret_cast<const int*>(dsc)[0], nover = reinterpret_cast<const int*>(dsc)[1];
    const InstDev* slist = reinterpret_cast<const InstDev*>(dsc + DL.inst);
    const int* over = reinterpret_cast<const int*>(dsc + DL.over);
    const double* flux_s = dyn_s + parity * dyn_len;
    const double* bkgfac = flux_s + kFoldLMax * P.smem_nT;
    const int fstride = P.smem_nT;
    const int n_sub = npx > kT64 ? 2 : 1;
    double stat = 0.0;

    for (int h = 0; h < n_sub; ++h) {
      const int p0h = p0 + kT64 * h;
      const int npxh = min(kT64, npx - kT64 * h);
      const bool last_sub = h + 1 == n_sub;
      // what comes after this sub-tile: the other half of the item, or the first half of the next item
      const bool more = !last_sub || next < n_items;
      const int n_item = last_sub ? next : item, n_h = last_sub ? 0 : 1;
      const bool expo_after = more && (last_sub ? nl_next > 0 : nl > 0);
      const uint32_t st = DB ? (ksub & 1u) : 0u, st_n = DB ? st ^ 1u : 0u;  // stage of this sub-tile / of the next one
      const float* es = reinterpret_cast<const float*>(es_base + st * st_e);
      const float* bs = reinterpret_cast<const float*>(bs_base + st * st_b);
      const float* cs = reinterpret_cast<const float*>(cs_base + st * st_c);
      const unsigned char* ms = ms_base + st * st_m;
      if (DB && more) {  // the other stage was drained before the last end-of-sub-tile barrier: request the next inputs now
        if (expo_after) issue_expo(n_item, n_h, st_n);
        issue_epi(n_item, n_h, st_n);
      }

      // this thread's pixel quad and, per overlapping source, a pointer into its cutout (zero word outside)
      const int pq = p0h + 4 * q;
      const bool qlive = 4 * q < npxh;
      const int qy = pq / nx, qx = pq - qy * nx;
      double acc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = 0.0;

      if (nl > 0) {
        const float* cptr[kFoldLMax];
        int pstr[kFoldLMax];
#pragma unroll
        for (int l = 0; l < kFoldLMax; ++l) {
          cptr[l] = P.zeros;
          pstr[l] = 0;
          if (l < nl && qlive) {
            const InstDev& in = slist[l];
            const int ly = qy - in.y0, lx = qx - in.x0a;
            if (ly >= 0 && ly < in.cy && lx >= 0 && lx < in.pitch) {
              cptr[l] = in.conv + ly * in.pitch + lx;
              pstr[l] = in.plane_stride;
            }
          }
        }
        mbar_wait(&bars[st], (eparity >> st) & 1u);  // exposure sub-tile has landed
        eparity ^= 1u << st;
        // ---- phase 1: v[t][p] = 1e4 * exposure[t,p] * sum_s F_s[t] * conv_s[t,p], all true bins in one pass ----
        if (nl <= 2) {
#pragma unroll
          constexpr int RB = DB ? 4 : 2;  // rows in flight (two loads each); four instead of two: no gain at three CTAs per SM
          for (int kb = 0; kb < kT64Rows; kb += RB) {
            float4 c0[RB], c1[RB];
#pragma unroll
            for (int u = 0; u < RB; ++u) {
              const int tc = min(slot + kT64Slots * (kb + u), nT - 1);
              c0[u] = ldg_f4_hint(cptr[0] + (size_t)tc * pstr[0], pol_keep);
              c1[u] = ldg_f4_hint(cptr[1] + (size_t)tc * pstr[1], pol_keep);
            }
#pragma unroll
            for (int u = 0; u < RB; ++u) {
              const int t = slot + kT64Slots * (kb + u);
              if (t < nT) {
                double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
                if (qlive) {
                  const double F0 = flux_s[t], F1 = flux_s[fstride + t];
                  s0 = fma(F1, (double)c1[u].x, fma(F0, (double)c0[u].x, s0));
                  s1 = fma(F1, (double)c1[u].y, fma(F0, (double)c0[u].y, s1));
                  s2 = fma(F1, (double)c1[u].z, fma(F0, (double)c0[u].z, s2));
                  s3 = fma(F1, (double)c1[u].w, fma(F0, (double)c0[u].w, s3));
                  const float4 e = *reinterpret_cast<const float4*>(es + t * kT64 + 4 * q);
                  s0 = (s0 * (double)e.x) * 1e4;
                  s1 = (s1 * (double)e.y) * 1e4;
                  s2 = (s2 * (double)e.z) * 1e4;
                  s3 = (s3 * (double)e.w) * 1e4;
                }
                double* dst = vs + t * kT64 + 4 * (GPY_T64_DMMA ? q ^ (t & 3) : q);  // (swizzled: see phase 2)
                *reinterpret_cast<double2*>(dst) = make_double2(s0, s1);
                *reinterpret_cast<double2*>(dst + 2) = make_double2(s2, s3);
              }
            }
          }
        } else {
#pragma unroll 1
          for (int k = 0; k < kT64Rows; ++k) {  // three or four sources: one row (four loads) in flight
            const int t = slot + kT64Slots * k;
            const int tc = min(t, nT - 1);
            float4 cg[kFoldLMax];
#pragma unroll
            for (int l = 0; l < kFoldLMax; ++l) cg[l] = ldg_f4_hint(cptr[l] + (size_t)tc * pstr[l], pol_keep);
            if (t < nT) {
              double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
              if (qlive) {
#pragma unroll
                for (int l = 0; l < kFoldLMax; ++l) {
                  const double F = flux_s[l * fstride + t];
                  s0 = fma(F, (double)cg[l].x, s0);
                  s1 = fma(F, (double)cg[l].y, s1);
                  s2 = fma(F, (double)cg[l].z, s2);
                  s3 = fma(F, (double)cg[l].w, s3);
                }
                // further sources on this tile (dense fields): one at a time
                for (int o = 0; o < nover; ++o) {
                  const InstDev& in = P.insts[over[o]];
                  const int ly = qy - in.y0, lx = qx - in.x0a;
                  if (ly < 0 || ly >= in.cy || lx < 0 || lx >= in.pitch) continue;
                  const float4 co = __ldg(reinterpret_cast<const float4*>(in.conv + ly * in.pitch + lx + (size_t)t * in.plane_stride));
                  const double F = __ldg(P.flux + in.flux_off + t);
                  s0 = fma(F, (double)co.x, s0);
                  s1 = fma(F, (double)co.y, s1);
                  s2 = fma(F, (double)co.z, s2);
                  s3 = fma(F, (double)co.w, s3);
                }
                const float4 e = *reinterpret_cast<const float4*>(es + t * kT64 + 4 * q);
                s0 = (s0 * (double)e.x) * 1e4;
                s1 = (s1 * (double)e.y) * 1e4;
                s2 = (s2 * (double)e.z) * 1e4;
                s3 = (s3 * (double)e.w) * 1e4;
              }
              double* dst = vs + t * kT64 + 4 * (GPY_T64_DMMA ? q ^ (t & 3) : q);
              *reinterpret_cast<double2*>(dst) = make_double2(s0, s1);
              *reinterpret_cast<double2*>(dst + 2) = make_double2(s2, s3);
            }
          }
        }
        __syncthreads();
        // exposure sub-tile fully consumed: fetch the next one while phase 2 / the epilogue run
        if (!DB && expo_after) issue_expo(n_item, n_h, 0);
        // ---- phase 2: acc[r] += pdf[t][r] * v[t][p] over the non-zero band ----
#if GPY_T64_DMMA
        // Warp w owns the pixels 8 w .. 8 w + 7 and all reco chunks: per chunk a loop over the k-steps (four true bins)
        // of its band, two 64-bit shared loads and one DMMA per 256 FMAs.  A = v^T[8 px][4 t] (the quad index of v is
        // XOR-swizzled with t & 3 by phase 1, so the 16 lanes of a half-warp hit 16 different bank pairs),
        // B = pdf[4 t][8 r] from the band-compressed table (zero outside the chunk's band), D -> acc[2 c + i] =
        // npred[px = 8 w + lane / 4][r = 8 c + 2 (lane % 4) + i].
        {
          const int lane = tid & 31, wq = tid >> 5, g = lane >> 2, tg = lane & 3;
          const double* arow = vs + tg * kT64 + 4 * ((2 * wq + (g >> 2)) ^ tg) + (g & 3);
          // one loop over the k-steps of the union of the bands: the A fragment is loaded once per k-step, the chunks
          // whose band the k-step touches (warp-uniform test) each add one DMMA -- four independent accumulator chains
          int blo[4], bhi[4], boff[4];
          int t_lo = nT, t_hi = 0;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int4 bnd = *reinterpret_cast<const int4*>(band_s + c * 4);  // (zero beyond nchunks: empty band)
            blo[c] = bnd.x;
            bhi[c] = c < nchunks ? bnd.y : bnd.x;
            boff[c] = (bnd.z - bnd.x) * 8 + g;
            if (bhi[c] > blo[c]) {
              t_lo = min(t_lo, blo[c]);
              t_hi = max(t_hi, bhi[c]);
            }
          }
          for (int t0 = t_lo & ~3; t0 < t_hi; t0 += 4) {
            const int t = t0 + tg;
            const double a = t < nT ? arow[t0 * kT64] : 0.0;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              if (t0 + 4 > blo[c] && t0 < bhi[c]) {
                const double bv = (t >= blo[c] && t < bhi[c]) ? pdf_s[boff[c] + t * 8] : 0.0;
                dmma884(acc[2 * c], acc[2 * c + 1], a, bv);
              }
            }
          }
        }
#else
        if (cslot < nchunks) {
          const int4 bnd = *reinterpret_cast<const int4*>(band_s + cslot * 4);
          const double* vrow = vs + px;
          const double* prow = pdf_s + (bnd.z - bnd.x) * 8;
#if GPY_T64_P2_PAIR
          for (int t = bnd.x; t < bnd.y; ++t) {  // two pixels x four reco bins: three 128-bit loads per eight FMAs
            const double2 v = *reinterpret_cast<const double2*>(vrow + t * kT64);
            const double2 m0 = *reinterpret_cast<const double2*>(prow + t * 8 + rh);
            const double2 m1 = *reinterpret_cast<const double2*>(prow + t * 8 + rh + 2);
            acc[0] = fma(v.x, m0.x, acc[0]);
            acc[1] = fma(v.x, m0.y, acc[1]);
            acc[2] = fma(v.x, m1.x, acc[2]);
            acc[3] = fma(v.x, m1.y, acc[3]);
            acc[4] = fma(v.y, m0.x, acc[4]);
            acc[5] = fma(v.y, m0.y, acc[5]);
            acc[6] = fma(v.y, m1.x, acc[6]);
            acc[7] = fma(v.y, m1.y, acc[7]);
          }
#else
          for (int t = bnd.x; t < bnd.y; ++t) {
            const double v = vrow[t * kT64];
            const double2 m0 = *reinterpret_cast<const double2*>(prow + t * 8);
            const double2 m1 = *reinterpret_cast<const double2*>(prow + t * 8 + 2);
            const double2 m2 = *reinterpret_cast<const double2*>(prow + t * 8 + 4);
            const double2 m3 = *reinterpret_cast<const double2*>(prow + t * 8 + 6);
            acc[0] = fma(v, m0.x, acc[0]);
            acc[1] = fma(v, m0.y, acc[1]);
            acc[2] = fma(v, m1.x, acc[2]);
            acc[3] = fma(v, m1.y, acc[3]);
            acc[4] = fma(v, m2.x, acc[4]);
            acc[5] = fma(v, m2.y, acc[5]);
            acc[6] = fma(v, m3.x, acc[6]);
            acc[7] = fma(v, m3.y, acc[7]);
          }
#endif
        }
#endif
        __syncthreads();  // v is rewritten by the parking step
      } else if (!DB && expo_after) {
        issue_expo(n_item, n_h, 0);  // no source on this tile: its exposure was never requested, the buffer is free
      }

      // ---- epilogue: background, clip, parked; then the Cash terms dealt over all threads ----
      mbar_wait(&bars[2 + st], (sparity >> st) & 1u);
      sparity ^= 1u << st;
      double* mus = vs;
#if GPY_T64_DMMA
      {
        const int lane = tid & 31, pxd = 8 * (tid >> 5) + (lane >> 2), tg = lane & 3;
        if (pxd < npxh) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const int r = 8 * c + 2 * tg + i;
              if (r < nR) {
                double mu = acc[2 * c + i];
                if (P.include_background) {
                  if (has_bkg) mu += (double)bs[r * kT64 + pxd] * bkgfac[r];
                  if (mu < 0.0) mu = 0.0;  // npred_total.data[npred_total.data < 0.0] = 0 (map.py:787)
                }
                mus[r * kT64 + pxd] = mu;
              }
            }
          }
        }
      }
#elif GPY_T64_P2_PAIR
      if (cslot < nchunks && px < npxh) {  // (npxh is even: P % 16 == 0 on the tensor-map path)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int r = 8 * cslot + rh + j;
          if (r >= nR) break;
          double mu0 = acc[j], mu1 = acc[4 + j];
          if (P.include_background) {
            if (has_bkg) {
              const float2 bg = *reinterpret_cast<const float2*>(bs + r * kT64 + px);
              mu0 += (double)bg.x * bkgfac[r];
              mu1 += (double)bg.y * bkgfac[r];
            }
            if (mu0 < 0.0) mu0 = 0.0;  // npred_total.data[npred_total.data < 0.0] = 0 (map.py:787)
            if (mu1 < 0.0) mu1 = 0.0;
          }
          *reinterpret_cast<double2*>(mus + r * kT64 + px) = make_double2(mu0, mu1);
        }
      }
#else
      if (cslot < nchunks && px < npxh) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int r = 8 * cslot + j;
          if (r >= nR) break;
          double mu = acc[j];
          if (P.include_background) {
            if (has_bkg) mu += (double)bs[r * kT64 + px] * bkgfac[r];
            if (mu < 0.0) mu = 0.0;  // npred_total.data[npred_total.data < 0.0] = 0 (map.py:787)
          }
          mus[r * kT64 + px] = mu;
        }
      }
#endif
      __syncthreads();
      const bool stage_next = last_sub && next < n_items;
      double dyn0 = 0.0, dyn1 = 0.0;
      if (stage_next) {
        mbar_wait(&bars[4 + (parity ^ 1u)], ((it + 1u) >> 1) & 1u);
        if (tid < dyn_len) dyn0 = load_dyn(parity ^ 1u, tid);
        if (tid + kFoldThreads < dyn_len) dyn1 = load_dyn(parity ^ 1u, tid + kFoldThreads);
      }
      {
        const double log_trunc = P.log_trunc;
        double stat_mu = 0.0, stat_nl = 0.0;
        for (int i = tid; i < nR * (kT64 / 2); i += kFoldThreads) {  // one row per warp and iteration: warp-uniform trip count
          const int r = i >> 5, lp = (i & 31) * 2;
          const bool live = lp < npxh;
          const double2 mu2 = *reinterpret_cast<const double2*>(mus + r * kT64 + lp);
          if (WRITE_NPRED && live)
            *reinterpret_cast<double2*>(P.npred_out + (size_t)r * (size_t)Ds.P + p0h + lp) = mu2;
          if (has_counts) {
            float2 n2;
            unsigned m2;
            if (!COMPACT) {
              n2 = *reinterpret_cast<const float2*>(cs + r * kT64 + lp);
              m2 = has_mask ? *reinterpret_cast<const unsigned short*>(ms + r * kT64 + lp) : 0x0101u;
            } else {
              const unsigned nn2 = *reinterpret_cast<const unsigned*>(reinterpret_cast<const unsigned short*>(cs) + r * kT64 + lp);
              n2 = make_float2((float)(nn2 & 0xffffu), (float)(nn2 >> 16));
              const unsigned mb = has_mask ? (unsigned)ms[r * 16 + ((lp + kT64 * h) >> 3)] >> (lp & 7) : 3u;
              m2 = (mb & 1u) | ((mb & 2u) << 7);
            }
            const bool on0 = live && (m2 & 0xffu), on1 = live && (m2 >> 8);
            const bool big0 = mu2.x > kTrunc, big1 = mu2.y > kTrunc;
            const double a0 = big0 ? mu2.x : kTrunc, a1 = big1 ? mu2.y : kTrunc;
            stat_mu += (on0 ? a0 : 0.0) + (on1 ? a1 : 0.0);
            const bool need0 = on0 && n2.x > 0.f, need1 = on1 && n2.y > 0.f;
            if (__any_sync(0xffffffffu, need0 || need1)) {
              double l0 = fast_log64(a0, logtab_s), l1 = fast_log64(a1, logtab_s);
              l0 = big0 ? (a0 <= 1.7976931348623157e308 ? l0 : a0) : log_trunc;
              l1 = big1 ? (a1 <= 1.7976931348623157e308 ? l1 : a1) : log_trunc;
              stat_nl += (need0 ? (double)n2.x * l0 : 0.0) + (need1 ? (double)n2.y * l1 : 0.0);
            }
          }
        }
        stat += stat_mu - stat_nl;
      }
      if (stage_next) {
        double* dst = dyn_s + (parity ^ 1u) * dyn_len;
        if (tid < dyn_len) dst[tid] = dyn0;
        if (tid + kFoldThreads < dyn_len) dst[tid + kFoldThreads] = dyn1;
        for (int i = tid + 2 * kFoldThreads; i < dyn_len; i += kFoldThreads) dst[i] = load_dyn(parity ^ 1u, i);
      }
      if (last_sub) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) stat += __shfl_xor_sync(0xffffffffu, stat, o);
        if ((tid & 31) == 0) wsum[tid >> 5] = stat;
        if (DYNAMIC && tid == kClaimTid) {
          const int c = gridDim.x + claim_raw;
          nxt_s = make_int2(c, item_at(c));
        }
      }
      __syncthreads();  // end of sub-tile: the epilogue inputs, mu and (last sub) the descriptor views are free
      if (!DB && more) issue_epi(n_item, n_h, 0);
      ++ksub;
    }
    if (tid == 0)
      P.partials[(size_t)b * P.n_tiles + tile] =
          ((wsum[0] + wsum[1]) + (wsum[2] + wsum[3])) + ((wsum[4] + wsum[5]) + (wsum[6] + wsum[7]));
    if (DYNAMIC) {
      const int2 v = nxt_s;
      wnn = v.x;
      nn = v.y;
    }
    const bool done = next >= n_items;
    item = next;
    w = wn;
    next = nn;
    wn = wnn;
    if (done) break;
  }
}

}  // namespace gpy
