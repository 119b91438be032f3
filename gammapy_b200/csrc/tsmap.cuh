// TS-map inner loop on sm_100a: per pixel a one-parameter Cash fit of `norm` in npred = background + norm * model,
// model = kernel x exposure on the kernel-sized cutout centred on the pixel.
//
// Reference (relative to /root/reference/gammapy):
//   estimators/map/ts.py:43-65      _extract_array                       cutout of the kernel's shape
//   estimators/map/ts.py:703-783    SimpleMapDataset (+ from_arrays)     model, invalid-bin removal, derivatives
//   estimators/map/ts.py:819-870    BrentqFluxEstimator.estimate_best_fit
//   estimators/map/ts.py:1053-1095  BrentqFluxEstimator.run              npred, npred_excess
//   stats/fit_statistics_cython.pyx:55-157  cash_sum / f_cash_root / norm_bounds
//   scipy.optimize.brentq (scipy/optimize/Zeros/brentq.c, not under /root/reference: restated, pinned against the
//   installed scipy in tests/test_oracle_tsmap.py)
//
// One warp per pixel.  Every pass over the cutout is a lane-strided loop followed by an xor-butterfly, so all lanes
// hold bit-identical sums and run the (scalar) Brent iteration redundantly without divergence.  Products and sums
// that the reference rounds separately use __dmul_rn / __dadd_rn so that FMA contraction cannot change them.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace gpy {

struct TsMapParams {
  const double* counts;      // [n_planes, ny, nx]
  const double* background;  // [n_planes, ny, nx]
  const double* exposure;    // [n_planes, ny, nx]
  const double* kernel;      // [n_planes, ky, kx]
  const unsigned char* mask; // [ny, nx] or null: positions to fit
  double* out;               // [kTsPlanes, ny, nx]
  int n_planes, ny, nx, ky, kx;
  int max_niter;
  double rtol, xtol, n_sigma;
  double log_trunc;
};

constexpr int kTsPlanes = 9;  // norm, norm_err, niter, ts, stat, stat_null, success, npred, npred_excess

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// The bins of one pixel's cutout in the reference's order (plane, dy, dx), dealt to the lanes round-robin.
struct TsWindow {
  const TsMapParams& P;
  int j, i, lane, n_bins, kk;
  __device__ TsWindow(const TsMapParams& p, int j_, int i_, int lane_)
      : P(p), j(j_), i(i_), lane(lane_), n_bins(p.n_planes * p.ky * p.kx), kk(p.ky * p.kx) {}

  // c, b, m of bin k; false when the bin is dropped (outside the map, or counts == background == model == 0)
  __device__ __forceinline__ bool load(int e, int dy, int dx, double& c, double& b, double& m) const {
    const int y = j + dy - P.ky / 2, x = i + dx - P.kx / 2;
    if (y < 0 || y >= P.ny || x < 0 || x >= P.nx) return false;
    const size_t o = ((size_t)e * P.ny + y) * P.nx + x;
    c = __ldg(P.counts + o);
    b = __ldg(P.background + o);
    m = __dmul_rn(__ldg(P.kernel + ((size_t)e * P.ky + dy) * P.kx + dx), __ldg(P.exposure + o));
    return !(c == 0.0 && b == 0.0 && m == 0.0);
  }

  template <typename F>
  __device__ __forceinline__ void for_each(F&& body) const {
    int e = 0, dy = 0, dx = lane;
    for (int k = lane; k < n_bins; k += 32) {
      while (dx >= P.kx) {
        dx -= P.kx;
        if (++dy == P.ky) {
          dy = 0;
          ++e;
        }
      }
      double c, b, m;
      if (load(e, dy, dx, c, b, m)) body(k, c, b, m);
      dx += 32;
    }
  }
};

// The same bins read from the shared-memory copy of a strip's union window (ts_map_strip_kernel): rows of
// `stride` = kx + kTsStrip - 1 doubles, planes and rows contiguous, zeros where the map ends (such bins drop out by
// the all-zero rule exactly like the reference's padded maps, ts.py:772-777).
constexpr int kTsStrip = 8;  // pixels (= warps) per block

struct TsWindowSmem {
  const double *c, *b, *e, *kern;  // shared memory
  int lane, n_bins, kx, stride, shift;
  template <typename F>
  __device__ __forceinline__ void for_each(F&& body) const {
    int dx = lane, off = lane + shift;
    for (int k = lane; k < n_bins; k += 32) {
      while (dx >= kx) {
        dx -= kx;
        off += stride - kx;
      }
      const double cc = c[off], bb = b[off];
      const double m = __dmul_rn(kern[k], e[off]);
      if (!(cc == 0.0 && bb == 0.0 && m == 0.0)) body(k, cc, bb, m);
      dx += 32;
      off += 32;
    }
  }
};

// f_cash_root_cython (.pyx:90-121): 2 sum_{m > 0} (c > 0 ? m (1 - c / (x m + b)) : m)
template <class Window>
__device__ __forceinline__ double ts_derivative(const Window& W, double x) {
  double s = 0.0;
  W.for_each([&](int, double c, double b, double m) {
    if (m > 0.0) {
      if (c > 0.0)
        s += __dmul_rn(m, __dadd_rn(1.0, -__ddiv_rn(c, __dadd_rn(__dmul_rn(x, m), b))));
      else
        s += m;
    }
  });
  return 2.0 * warp_sum(s);
}

__device__ __forceinline__ double cash_term(double c, double npred, double log_trunc) {
  double lognpr = 0.0;
  if (npred > kTrunc) {
    if (c > 0.0) lognpr = log(npred);
  } else {
    npred = kTrunc;
    lognpr = log_trunc;
  }
  double t = npred;
  if (c > 0.0) t -= c * lognpr;
  return t;
}

__device__ __forceinline__ bool signbit_d(double v) { return __double2hiint(v) < 0; }

// The fit of one pixel by one warp; W deals the bins of its cutout to the lanes.
template <class Window>
__device__ __forceinline__ void ts_fit_pixel(const TsMapParams& P, const Window& W, long long pix, int lane) {
  const long long n_pix = (long long)P.ny * P.nx;

  // ---- norm_bounds_cython (.pyx:124-157) + the two sums estimate_best_fit tests (ts.py:836)
  double s_model = 0.0, s_counts = 0.0, sum_counts = 0.0, sum_model = 0.0, sum_bkg = 0.0;
  double sn_min = 1e14, c_min = 1.0, sn_min_total = 1e14;
  int k_min = 0x7fffffff;
  W.for_each([&](int k, double c, double b, double m) {
    sum_counts += c;
    sum_model += m;
    sum_bkg += b;
    if (c > 0.0) {
      s_counts += c;
      if (m > 0.0) {
        const double sn = __ddiv_rn(b, m);
        if (sn < sn_min) {
          sn_min = sn;
          c_min = c;
          k_min = k;
        }
      }
    }
    if (m > 0.0) {
      s_model += m;
      const double sn = __ddiv_rn(b, m);
      if (sn < sn_min_total) sn_min_total = sn;
    }
  });
  s_model = warp_sum(s_model);
  s_counts = warp_sum(s_counts);
  sum_counts = warp_sum(sum_counts);
  sum_model = warp_sum(sum_model);
  sum_bkg = warp_sum(sum_bkg);
  sn_min_total = warp_min(sn_min_total);
  {  // strict "<" in index order: the first bin that attains the minimum supplies c_min
    const double best = warp_min(sn_min);
    int k = sn_min == best ? k_min : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) k = min(k, __shfl_xor_sync(0xffffffffu, k, o));
    const unsigned owner = __ballot_sync(0xffffffffu, sn_min == best && k_min == k);
    c_min = __shfl_sync(0xffffffffu, c_min, owner ? __ffs(owner) - 1 : 0);
    if (!owner) c_min = 1.0;
    sn_min = best;
  }
  const double b_min = __dadd_rn(__ddiv_rn(c_min, s_model), -sn_min);
  const double b_max = __dadd_rn(__ddiv_rn(s_counts, s_model), -sn_min);
  const double norm_min_total = -sn_min_total;

  // ---- estimate_best_fit (ts.py:819-858): Brent root of the derivative on [b_min, b_max]
  double norm = norm_min_total;
  int niter = 0;
  bool success = true;
  if (!(sum_counts <= 0.0 || sum_model <= 0.0)) {
    // scipy.optimize.brentq(f, a, b, xtol=2e-12, rtol, maxiter, full_output=True, disp=True)
    double xpre = b_min, xcur = b_max, xblk = 0.0, fblk = 0.0, spre = 0.0, scur = 0.0;
    double fpre = ts_derivative(W, xpre), fcur = ts_derivative(W, xcur);
    double root = 0.0;
    int status = 0;  // 0 running, 1 converged, -1 sign error / nan, -2 not converged
    if (!(xpre == xpre) || !(xcur == xcur) || !(fpre == fpre) || !(fcur == fcur)) {
      status = -1;  // scipy raises ValueError for nan bounds / RuntimeError for nan values: both are caught (ts.py:855)
    } else if (fpre == 0.0) {
      root = xpre;
      status = 1;
    } else if (fcur == 0.0) {
      root = xcur;
      status = 1;
    } else if (signbit_d(fpre) == signbit_d(fcur)) {
      status = -1;
    }
    if (status == 0) {
      status = -2;
      for (int it = 0; it < P.max_niter; ++it) {
        ++niter;
        if (fpre != 0.0 && fcur != 0.0 && signbit_d(fpre) != signbit_d(fcur)) {
          xblk = xpre;
          fblk = fpre;
          spre = scur = xcur - xpre;
        }
        if (fabs(fblk) < fabs(fcur)) {
          xpre = xcur;
          xcur = xblk;
          xblk = xpre;
          fpre = fcur;
          fcur = fblk;
          fblk = fpre;
        }
        const double delta = __dmul_rn(__dadd_rn(P.xtol, __dmul_rn(P.rtol, fabs(xcur))), 0.5);
        const double sbis = __dmul_rn(__dadd_rn(xblk, -xcur), 0.5);
        if (fcur == 0.0 || fabs(sbis) < delta) {
          root = xcur;
          status = 1;
          break;
        }
        if (fabs(spre) > delta && fabs(fcur) < fabs(fpre)) {
          double stry;
          if (xpre == xblk) {  // interpolate
            stry = __ddiv_rn(__dmul_rn(-fcur, __dadd_rn(xcur, -xpre)), __dadd_rn(fcur, -fpre));
          } else {  // extrapolate
            const double dpre = __ddiv_rn(__dadd_rn(fpre, -fcur), __dadd_rn(xpre, -xcur));
            const double dblk = __ddiv_rn(__dadd_rn(fblk, -fcur), __dadd_rn(xblk, -xcur));
            stry = __ddiv_rn(__dmul_rn(-fcur, __dadd_rn(__dmul_rn(fblk, dblk), -__dmul_rn(fpre, dpre))),
                             __dmul_rn(__dmul_rn(dblk, dpre), __dadd_rn(fblk, -fpre)));
          }
          if (__dmul_rn(2.0, fabs(stry)) < fmin(fabs(spre), __dadd_rn(__dmul_rn(3.0, fabs(sbis)), -delta))) {
            spre = scur;  // good short step
            scur = stry;
          } else {
            spre = sbis;  // bisect
            scur = sbis;
          }
        } else {
          spre = sbis;
          scur = sbis;
        }
        xpre = xcur;
        fpre = fcur;
        if (fabs(scur) > delta)
          xcur += scur;
        else
          xcur += (sbis > 0.0 ? delta : -delta);
        fcur = ts_derivative(W, xcur);
        if (!(fcur == fcur)) {  // scipy: "The function value at x=... is NaN; solver cannot continue."
          status = -1;
          break;
        }
      }
    }
    if (status == 1) {
      norm = fmax(root, norm_min_total);
    } else {  // except (RuntimeError, ValueError): norm_min_total, max_niter, False (ts.py:855-856)
      norm = norm_min_total;
      niter = P.max_niter;
      success = false;
    }
  }

  // ---- second derivative, stat at the fitted norm and at norm = 0, npred (ts.py:733-746, 858-870, 1090-1093)
  double d2 = 0.0, stat = 0.0, stat_null = 0.0, npred = 0.0;
  W.for_each([&](int, double c, double b, double m) {
    const double mu = __dadd_rn(b, __dmul_rn(norm, m));
    const double bottom = __dmul_rn(mu, mu);
    if (bottom != 0.0) d2 += __ddiv_rn(__dmul_rn(__dmul_rn(m, m), c), bottom);
    stat += cash_term(c, mu, P.log_trunc);
    stat_null += cash_term(c, __dadd_rn(b, __dmul_rn(0.0, m)), P.log_trunc);
    npred += mu;
  });
  d2 = warp_sum(d2);
  stat = 2.0 * warp_sum(stat);
  stat_null = 2.0 * warp_sum(stat_null);
  npred = warp_sum(npred);
  if (lane == 0) {
    double* o = P.out + pix;
    o[0 * n_pix] = norm;
    o[1 * n_pix] = sqrt(1.0 / d2) * P.n_sigma;
    o[2 * n_pix] = (double)niter;
    o[3 * n_pix] = stat_null - stat;
    o[4 * n_pix] = stat;
    o[5 * n_pix] = stat_null;
    o[6 * n_pix] = success ? 1.0 : 0.0;
    o[7 * n_pix] = npred;
    o[8 * n_pix] = npred - sum_bkg;
  }
}

// Any shape: every pass reads the maps through L1 / L2.
__global__ void __launch_bounds__(256) ts_map_kernel(const TsMapParams P) {
  const int lane = threadIdx.x & 31;
  const long long pix = ((long long)blockIdx.x * 256 + threadIdx.x) >> 5;
  const long long n_pix = (long long)P.ny * P.nx;
  if (pix >= n_pix) return;
  if (P.mask && !P.mask[pix]) {
    if (lane < kTsPlanes) P.out[(size_t)lane * n_pix + pix] = __longlong_as_double(0x7ff8000000000000ll);
    return;
  }
  const TsWindow W(P, (int)(pix / P.nx), (int)(pix % P.nx), lane);
  ts_fit_pixel(P, W, pix, lane);
}

// Shared-memory variant: a block fits kTsStrip consecutive pixels of one row (one per warp).  The union of their
// cutouts ([n_planes][ky][kx + kTsStrip - 1] of counts, background, exposure) and the kernel are staged once and
// every pass of every warp reads shared memory only.  Dynamic shared memory: ts_strip_smem_bytes().
__host__ __device__ inline size_t ts_strip_smem_bytes(int n_planes, int ky, int kx) {
  return ((size_t)3 * n_planes * ky * (kx + kTsStrip - 1) + (size_t)n_planes * ky * kx) * sizeof(double);
}

__global__ void __launch_bounds__(32 * kTsStrip) ts_map_strip_kernel(const TsMapParams P) {
  extern __shared__ double ts_smem[];
  const int strips = (P.nx + kTsStrip - 1) / kTsStrip;
  const int j = blockIdx.x / strips, i0 = (blockIdx.x % strips) * kTsStrip;
  const int stride = P.kx + kTsStrip - 1, rows = P.n_planes * P.ky;
  const int plane = rows * stride;
  double* sc = ts_smem;
  double* sb = sc + plane;
  double* se = sb + plane;
  double* sk = se + plane;
  for (int w = threadIdx.x; w < plane; w += blockDim.x) {
    const int row = w / stride, u = w - row * stride;
    const int e = row / P.ky, dy = row - e * P.ky;
    const int y = j + dy - P.ky / 2, x = i0 + u - P.kx / 2;
    double c = 0.0, b = 0.0, ex = 0.0;
    if (y >= 0 && y < P.ny && x >= 0 && x < P.nx) {
      const size_t o = ((size_t)e * P.ny + y) * P.nx + x;
      c = __ldg(P.counts + o);
      b = __ldg(P.background + o);
      ex = __ldg(P.exposure + o);
    }
    sc[w] = c;
    sb[w] = b;
    se[w] = ex;
  }
  for (int w = threadIdx.x; w < rows * P.kx; w += blockDim.x) sk[w] = __ldg(P.kernel + w);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i = i0 + warp;
  if (i >= P.nx) return;
  const long long n_pix = (long long)P.ny * P.nx, pix = (long long)j * P.nx + i;
  if (P.mask && !P.mask[pix]) {
    if (lane < kTsPlanes) P.out[(size_t)lane * n_pix + pix] = __longlong_as_double(0x7ff8000000000000ll);
    return;
  }
  TsWindowSmem W;
  W.c = sc;
  W.b = sb;
  W.e = se;
  W.kern = sk;
  W.lane = lane;
  W.n_bins = rows * P.kx;
  W.kx = P.kx;
  W.stride = stride;
  W.shift = warp;
  ts_fit_pixel(P, W, pix, lane);
}

// ---- compiled-stat registry entries over caller-provided 1-D arrays (utils/compilation.py:35-87) -----------------
// f_cash_root_compiled(x, counts, background, model) -> out[0]; norm_bounds_compiled(counts, background, model) ->
// out[0..2] = (b_min, b_max, -sn_min_total).  One block; fixed reduction order.
__global__ void __launch_bounds__(256)
f_cash_root_kernel(double x, const double* __restrict__ counts, const double* __restrict__ background,
                   const double* __restrict__ model, long long n, double* __restrict__ out) {
  __shared__ double sm[256];
  double s = 0.0;
  for (long long i = threadIdx.x; i < n; i += 256) {
    const double m = model[i], c = counts[i];
    if (m > 0.0) {
      if (c > 0.0)
        s += __dmul_rn(m, __dadd_rn(1.0, -__ddiv_rn(c, __dadd_rn(__dmul_rn(x, m), background[i]))));
      else
        s += m;
    }
  }
  sm[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = 2.0 * sm[0];
}

__global__ void __launch_bounds__(256)
norm_bounds_kernel(const double* __restrict__ counts, const double* __restrict__ background,
                   const double* __restrict__ model, long long n, double* __restrict__ out) {
  __shared__ double s_model[256], s_counts[256], sn_min[256], c_min[256], sn_tot[256];
  __shared__ long long k_min[256];
  double a_model = 0.0, a_counts = 0.0, a_sn = 1e14, a_c = 1.0, a_tot = 1e14;
  long long a_k = 0x7fffffffffffffffll;
  for (long long i = threadIdx.x; i < n; i += 256) {
    const double c = counts[i], m = model[i];
    if (c > 0.0) {
      a_counts += c;
      if (m > 0.0) {
        const double sn = __ddiv_rn(background[i], m);
        if (sn < a_sn) {
          a_sn = sn;
          a_c = c;
          a_k = i;
        }
      }
    }
    if (m > 0.0) {
      a_model += m;
      const double sn = __ddiv_rn(background[i], m);
      if (sn < a_tot) a_tot = sn;
    }
  }
  const int t = threadIdx.x;
  s_model[t] = a_model;
  s_counts[t] = a_counts;
  sn_min[t] = a_sn;
  c_min[t] = a_c;
  k_min[t] = a_k;
  sn_tot[t] = a_tot;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (t < o) {
      s_model[t] += s_model[t + o];
      s_counts[t] += s_counts[t + o];
      sn_tot[t] = fmin(sn_tot[t], sn_tot[t + o]);
      // strict "<" in index order: smaller value wins, equal values keep the smaller index
      if (sn_min[t + o] < sn_min[t] || (sn_min[t + o] == sn_min[t] && k_min[t + o] < k_min[t])) {
        sn_min[t] = sn_min[t + o];
        c_min[t] = c_min[t + o];
        k_min[t] = k_min[t + o];
      }
    }
    __syncthreads();
  }
  if (t == 0) {
    out[0] = __dadd_rn(__ddiv_rn(c_min[0], s_model[0]), -sn_min[0]);
    out[1] = __dadd_rn(__ddiv_rn(s_counts[0], s_model[0]), -sn_min[0]);
    out[2] = -sn_tot[0];
  }
}

}  // namespace gpy
