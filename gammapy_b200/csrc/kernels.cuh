// kernels.cuh -- sm_100a device code of libgpy_b200: the forward-fold likelihood hot path.
//
// Kernels (numbers follow BASELINE.json north_star / DESIGN.md):
//   K0  spectral_flux_kernel   energy-bin integrals of the spectral models           (a9 in SURVEY.md §8)
//   K1  spatial_fill_kernel / spatial_oversample_kernel / solid_angle_kernel         (a10, a10p, a10g, a10d)
//   K2  psf_conv_kernel        per-energy direct stencil, smem tiled, fp32            (a11)
//   K3+K4 fold_cash_tma_kernel (bulk-copy staged, persistent) / fold_cash_generic_kernel (any shape)
//                              exposure x flux -> edisp contraction -> + background -> masked Cash,
//                              fp64 accumulate, batch of parameter vectors per launch  (a2-a6, a12-a14)
//   prep_items_kernel          per-tile work descriptors (overlap lists); cached by the host across steps
//   dedup_*_kernel             batched calls: worklist of the (tile, vector) items that differ from vector 0
//   reduce_stat_kernel         deterministic second stage of the stat reduction
//   cash_sum_kernel            stand-alone (weighted) Cash for the compiled-stat registry (a3, a3')
// The TS-map kernels and the other two registry entries live in tsmap.cuh.
//
// Reference citations are relative to /root/reference/gammapy.
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <math.h>
#include <stdint.h>

namespace gpy {

constexpr double kDeg = 0.017453292519943295;  // astropy deg -> rad factor (pi / 180)
constexpr double kTrunc = 1e-25;               // stats/fit_statistics_cython.pyx:12-13
constexpr double kEdge95 = 2.326174307353347;  // modeling/models/spatial.py:974

// ----------------------------------------------------------------------------------------------
// WCS: CAR about the equator.  lon = crval + cdelt_x (i + 1 - crpix_x), lat = cdelt_y (j + 1 - crpix_y)
// (maps/wcs/geom.py:401-413; wcslib linp2x simple path).  No FMA contraction: keep numpy's rounding.
// CAR about a reference point off the equator (WcsGeom.create sets CRVAL = skydir, geom.py:370-410) is an oblique
// plate carree: native (phi, theta) = (x, y), the native origin sits at the reference point and -- with the default
// LONPOLE (0 deg for delta_0 >= 0, else 180 deg) and LATPOLE (+90 deg), FITS WCS paper II section 2.4 and eqs. 8-10 --
// the native pole lies on the reference meridian, so the spherical rotation (eq. 2) is a tilt of the native equator's
// origin up to latitude delta_0: rotate the unit vector about the y axis by delta_0, then add alpha_0.
// TAN about any reference point: gnomonic, native pole at the reference point.
// ----------------------------------------------------------------------------------------------
struct WcsDev {
  double crval_lon, binsz, crpix_x, crpix_y;
  double crval_lat;
  int proj;  // 0 CAR about (crval_lon, crval_lat) (crval_lat == 0: the separable equatorial form), 1 TAN
};

__device__ __forceinline__ void pix_to_world_deg(const WcsDev& w, double px, double py, double& lon,
                                                 double& lat) {
  if (w.proj == 1) {
    // TAN (FITS WCS paper II, Calabretta & Greisen 2002): R = (180 / pi) cot(theta), x = R sin(phi), y = -R cos(phi)
    // (eqs. 54, 14, 15), native pole at the reference point, LONPOLE = 180 deg, spherical rotation eq. (2)
    const double x = -w.binsz * (px + 1.0 - w.crpix_x), y = w.binsz * (py + 1.0 - w.crpix_y);
    const double r = hypot(x, y);
    if (r == 0.0) {
      lon = w.crval_lon;
      lat = w.crval_lat;
      return;
    }
    const double phi = atan2(x, -y), theta = atan2(180.0 / 3.141592653589793, r);
    double st, ct, sp, cp, sd, cd;
    sincos(theta, &st, &ct);
    sincos(w.crval_lat * kDeg, &sp, &cp);
    sincos(phi - 3.141592653589793, &sd, &cd);
    lat = asin(fmin(fmax(st * sp + ct * cp * cd, -1.0), 1.0)) / kDeg;
    lon = w.crval_lon + atan2(-ct * sd, st * cp - ct * sp * cd) / kDeg;
    return;
  }
  lon = __dadd_rn(w.crval_lon, __dmul_rn(-w.binsz, __dadd_rn(__dadd_rn(px, 1.0), -w.crpix_x)));
  lat = __dmul_rn(w.binsz, __dadd_rn(__dadd_rn(py, 1.0), -w.crpix_y));
  if (w.crval_lat != 0.0) {  // oblique plate carree: (lon - crval_lon, lat) so far are the native (phi, theta)
    double sp, cp, st, ct, s0, c0;
    sincos((lon - w.crval_lon) * kDeg, &sp, &cp);
    sincos(lat * kDeg, &st, &ct);
    sincos(w.crval_lat * kDeg, &s0, &c0);
    const double x = ct * cp * c0 - st * s0, y = ct * sp, z = ct * cp * s0 + st * c0;
    lon = w.crval_lon + atan2(y, x) / kDeg;
    lat = atan2(z, hypot(x, y)) / kDeg;
  }
}

// astropy.coordinates.angular_separation (Vincenty), radians.
__device__ __forceinline__ double angular_separation(double lon1, double lat1, double lon2, double lat2) {
  double sdlon, cdlon, slat1, clat1, slat2, clat2;
  sincos(lon2 - lon1, &sdlon, &cdlon);
  sincos(lat1, &slat1, &clat1);
  sincos(lat2, &slat2, &clat2);
  double num1 = clat2 * sdlon;
  double num2 = clat1 * slat2 - slat1 * clat2 * cdlon;
  double den = slat1 * slat2 + clat1 * clat2 * cdlon;
  return atan2(hypot(num1, num2), den);
}

// astropy.coordinates.position_angle, radians.
__device__ __forceinline__ double position_angle(double lon1, double lat1, double lon2, double lat2) {
  double sdl, cdl, s1, c1, s2, c2;
  sincos(lon2 - lon1, &sdl, &cdl);
  sincos(lat1, &s1, &c1);
  sincos(lat2, &s2, &c2);
  double x = s2 * c1 - c2 * s1 * cdl;
  double y = sdl * c2;
  return atan2(y, x);
}

// ----------------------------------------------------------------------------------------------
// WcsGeom.solid_angle (maps/wcs/geom.py:816-854): planar-triangle area from the four pixel corners.
// ----------------------------------------------------------------------------------------------
__global__ void solid_angle_kernel(WcsDev w, int nx, int ny, double* __restrict__ out) {
  int x = blockIdx.x * blockDim.x + threadIdx.x;
  int y = blockIdx.y;
  if (x >= nx || y >= ny) return;
  // corner grid coord[yy][xx], yy in {y, y+1} (edge y-0.5, y+0.5), xx likewise.
  double lon[2][2], lat[2][2];
#pragma unroll
  for (int iy = 0; iy < 2; ++iy)
#pragma unroll
    for (int ix = 0; ix < 2; ++ix) {
      double lo, la;
      pix_to_world_deg(w, (double)x - 0.5 + ix, (double)y - 0.5 + iy, lo, la);
      lon[iy][ix] = lo * kDeg;
      lat[iy][ix] = la * kDeg;
    }
  // names as in the reference: first index is the y edge, second the x edge
  // low_left = [:-1,:-1] -> (0,0); low_right = [1:,:-1] -> (1,0); up_left = [:-1,1:] -> (0,1); up_right -> (1,1)
  double ll_lon = lon[0][0], ll_lat = lat[0][0];
  double lr_lon = lon[1][0], lr_lat = lat[1][0];
  double ul_lon = lon[0][1], ul_lat = lat[0][1];
  double ur_lon = lon[1][1], ur_lat = lat[1][1];
  double low = angular_separation(ll_lon, ll_lat, lr_lon, lr_lat);
  double left = angular_separation(ll_lon, ll_lat, ul_lon, ul_lat);
  double up = angular_separation(ul_lon, ul_lat, ur_lon, ur_lat);
  double right = angular_separation(lr_lon, lr_lat, ur_lon, ur_lat);
  double angle_low_right =
      position_angle(lr_lon, lr_lat, ur_lon, ur_lat) - position_angle(lr_lon, lr_lat, ll_lon, ll_lat);
  double angle_up_left =
      position_angle(ul_lon, ul_lat, ur_lon, ur_lat) - position_angle(ll_lon, ll_lat, ul_lon, ul_lat);
  double area_low_right = 0.5 * low * right * sin(angle_low_right);
  double area_up_left = 0.5 * up * left * sin(angle_up_left);
  out[(size_t)y * nx + x] = fabs(area_low_right + area_up_left);
}

// ----------------------------------------------------------------------------------------------
// K1: SpatialModel.integrate_geom (modeling/models/spatial.py:222-309, 609-631)
// ----------------------------------------------------------------------------------------------
struct SpatialJob {
  int type;            // gpy_spatial_type
  int f;               // oversampling factor
  int ecx, ecy;        // evaluator (exposure cutout) image size
  int pitch, xpad;     // output row pitch (floats) and left padding: out[y * pitch + xpad + x]
  int ix0, iy0, icx, icy;  // inner cutout (2 max(r_eval, pix)) in evaluator pixel coords, f > 1 only
  int ex0, ey0;        // evaluator origin in parent pixels (solid angle lookup)
  int parent_nx;
  int pad_;
  WcsDev wcs_e;        // evaluator geometry
  WcsDev wcs_up;       // upsampled inner cutout geometry (f > 1)
  double lon0, lat0;   // source position, rad
  double a, norm;      // gauss e == 0: a = 1 - cos(sigma); norm of gauss / disk
  double major, ecc, phi;  // rad; phi wrapped to [0, pi)
  double edge_width;
  double px0, py0;     // point source: position in evaluator pixel coordinates
  const double* solid_angle;  // parent [ny, nx] sr
  float* out;          // [ecy, pitch]
  // The model's frame differs from the map's (SpatialModel.evaluate_geom: geom.get_coord(frame=self.frame),
  // spatial.py:196-220): pixel coordinates are rotated into the model frame before the model is evaluated, so
  // lon0 / lat0 / phi keep their meaning.  rot = map frame -> model frame, row major.
  int rotate, pad2_;
  double rot[9];
};

__device__ __forceinline__ void to_model_frame(const SpatialJob& j, double& lon_deg, double& lat_deg) {
  if (!j.rotate) return;
  double sl, cl, sb, cb;
  sincos(lon_deg * kDeg, &sl, &cl);
  sincos(lat_deg * kDeg, &sb, &cb);
  const double x = cb * cl, y = cb * sl, z = sb;
  const double x2 = j.rot[0] * x + j.rot[1] * y + j.rot[2] * z;
  const double y2 = j.rot[3] * x + j.rot[4] * y + j.rot[5] * z;
  const double z2 = j.rot[6] * x + j.rot[7] * y + j.rot[8] * z;
  lon_deg = atan2(y2, x2) / kDeg;
  lat_deg = atan2(z2, hypot(x2, y2)) / kDeg;
}

__device__ __forceinline__ double sigma_eff_dev(const SpatialJob& j, double lon, double lat) {
  // compute_sigma_eff (spatial.py:57-67)
  double phi_0 = position_angle(j.lon0, j.lat0, lon, lat);
  double d_phi = j.phi - phi_0;
  double minor = j.major * sqrt(1.0 - j.ecc * j.ecc);
  double sd, cd;
  sincos(d_phi, &sd, &cd);
  double a2 = (j.major * sd) * (j.major * sd);
  double b2 = (minor * cd) * (minor * cd);
  return j.major * minor / sqrt(a2 + b2);
}

__device__ __forceinline__ double eval_spatial(const SpatialJob& j, double lon_deg, double lat_deg) {
  double lon = lon_deg * kDeg, lat = lat_deg * kDeg;
  double sep = angular_separation(lon, lat, j.lon0, j.lat0);
  if (j.type == 1) {  // GaussianSpatialModel.evaluate (spatial.py:684-701)
    double a = j.a;
    if (j.ecc != 0.0) a = 1.0 - cos(sigma_eff_dev(j, lon, lat));
    double exponent = -0.5 * ((1.0 - cos(sep)) / a);
    return j.norm * exp(exponent);
  }
  // DiskSpatialModel.evaluate (spatial.py:977-993)
  double sigma_eff = j.major;
  if (j.ecc != 0.0) sigma_eff = sigma_eff_dev(j, lon, lat);
  double value = (sep - sigma_eff) / (sigma_eff * j.edge_width);
  return j.norm * (0.5 * (1.0 - erf(value * kEdge95)));
}

// One thread per output element (padded pitch): zero padding, point weights, f == 1 evaluation.
__device__ __forceinline__ void spatial_fill_body(const SpatialJob& j, int bx, int y) {
  int xp = bx * blockDim.x + threadIdx.x;
  if (xp >= j.pitch || y >= j.ecy) return;
  int x = xp - j.xpad;
  float v = 0.f;
  if (x >= 0 && x < j.ecx) {
    if (j.type == 0) {
      // PointSpatialModel._grid_weights (spatial.py:589-598): no solid angle
      double dx = fabs((double)x - j.px0);
      dx = dx < 1.0 ? 1.0 - dx : 0.0;
      double dy = fabs((double)y - j.py0);
      dy = dy < 1.0 ? 1.0 - dy : 0.0;
      v = (float)(dx * dy);
    } else if (j.f == 1) {
      double lon, lat;
      pix_to_world_deg(j.wcs_e, (double)x, (double)y, lon, lat);
      to_model_frame(j, lon, lat);
      double g = eval_spatial(j, lon, lat);
      double omega = j.solid_angle[(size_t)(j.ey0 + y) * j.parent_nx + (j.ex0 + x)];
      v = (float)(g * omega);  // float64 result, cast at the convolution input (ndmap.py:984)
    }
  }
  j.out[(size_t)y * j.pitch + xp] = v;
}
__global__ void spatial_fill_kernel(SpatialJob j) { spatial_fill_body(j, blockIdx.x, blockIdx.y); }

// f > 1: one warp per inner-cutout pixel, lanes stride the f*f sub-pixels; the block sum is rounded
// to float32 by the stack into the float32 map (spatial.py:295-296, ndmap.py:1170), then * solid angle.
__device__ __forceinline__ void spatial_oversample_body(const SpatialJob& j, int bx) {
  int warp = (bx * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (warp >= j.icx * j.icy) return;
  int ix = warp % j.icx, iy = warp / j.icx;
  int f = j.f, ff = f * f;
  double sum = 0.0;
  for (int k = lane; k < ff; k += 32) {
    int sx = k % f, sy = k / f;
    double lon, lat;
    pix_to_world_deg(j.wcs_up, (double)(ix * f + sx), (double)(iy * f + sy), lon, lat);
    to_model_frame(j, lon, lat);
    double g = eval_spatial(j, lon, lat);
    // numpy: evaluate(...) / f**2 is a true division per element
    g = __ddiv_rn(g, (double)ff);
    if (!isnan(g)) sum += g;  // block_reduce(func=np.nansum)
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (lane == 0) {
    int x = j.ix0 + ix, y = j.iy0 + iy;
    float s32 = (float)sum;
    if (!isfinite(s32)) s32 = 0.f;  // stack(nan_to_num=True)
    double omega = j.solid_angle[(size_t)(j.ey0 + y) * j.parent_nx + (j.ex0 + x)];
    j.out[(size_t)y * j.pitch + j.xpad + x] = (float)((double)s32 * omega);
  }
}
__global__ void spatial_oversample_kernel(SpatialJob j) { spatial_oversample_body(j, blockIdx.x); }

// ----------------------------------------------------------------------------------------------
// K2: WcsNDMap.convolve(PSFKernel), mode "same" (maps/wcs/ndmap.py:920-1010) as a direct stencil.
//   image  [cy][pitch] float32, zero outside [xpad, xpad + cx)
//   planes: per energy the kernel flipped, cropped to its non-zero half width h and padded to kw
//   out    [plane][cy][pitch] float32 (zero in the padding columns)
// Tile 64 x 32 outputs per 256-thread block; each thread 8 outputs (two float4 groups 32 apart, so
// the LDS.128 of a quarter warp are contiguous); fp32 FMA along a kernel row, fp64 across rows.
// ----------------------------------------------------------------------------------------------
struct ConvPlane {
  int h;     // tight half width of the plane's support
  int kw;    // padded row width of the flipped kernel (multiple of 4, >= 2h + 1)
  int koff;  // offset (floats) of the plane in the flipped-kernel buffer
  int pad_;
};

constexpr int kConvTX = 64, kConvTY = 32, kConvThreads = 256;

__device__ __forceinline__ void
psf_conv_body(const float* __restrict__ image, int cy, int cx, int pitch, int xpad,
              const float* __restrict__ kflip, const ConvPlane* __restrict__ planes,
              float* __restrict__ out, size_t plane_stride, int row0, int row1, int bx, int by, int bz) {
  // [row0, row1): rows of `image` that can hold a non-zero (a point source fills two rows, an oversampled model its
  // inner cutout).  Stencil rows that only see zeros are skipped -- they would add 0 * k = 0 to every sum --, so a
  // point source costs 2 of the 2 h + 1 rows and output tiles beyond the kernel's reach are written as zeros at once.
  extern __shared__ __align__(16) float conv_smem[];
  const ConvPlane pl = planes[bz];
  const int h = pl.h, kw = pl.kw, nrow = 2 * h + 1;
  const int spitch = kConvTX + kw;        // multiple of 4
  const int srows = kConvTY + 2 * h;
  float* in_s = conv_smem;                // [srows][spitch]
  float* k_s = conv_smem + (size_t)srows * spitch;  // [nrow][kw]
  const int tx0 = bx * kConvTX, ty0 = by * kConvTY;
  const int tid = threadIdx.x;
  if (ty0 - h >= row1 || ty0 + kConvTY + h <= row0) {  // block uniform: no input row of this tile can be non-zero
    const int y = ty0 + (tid >> 3);
    if (y < cy) {
      float* o = out + bz * plane_stride + (size_t)y * pitch + tx0;
      for (int k = (tid & 7); k < kConvTX && tx0 + k < pitch; k += 8) o[k] = 0.f;
    }
    return;
  }

  for (int i = tid; i < nrow * kw; i += kConvThreads) k_s[i] = kflip[pl.koff + i];
  for (int i = tid; i < srows * spitch; i += kConvThreads) {
    int ry = i / spitch, c = i - ry * spitch;
    int y = ty0 - h + ry, xp = tx0 - h + c;
    float v = 0.f;
    if (y >= 0 && y < cy && xp >= 0 && xp < pitch) v = image[(size_t)y * pitch + xp];
    in_s[i] = v;
  }
  __syncthreads();

  const int tx = tid & 7, ly = tid >> 3;
  const int xa = 4 * tx, xb = 32 + 4 * tx;
  double acc_a[4] = {0, 0, 0, 0}, acc_b[4] = {0, 0, 0, 0};
  // stencil row i reads image row ty0 - h + ly + i
  const int i_lo = max(0, row0 - (ty0 - h + ly)), i_hi = min(nrow, row1 - (ty0 - h + ly));
  for (int i = i_lo; i < i_hi; ++i) {
    const float* r = in_s + (size_t)(ly + i) * spitch;
    const float* kr = k_s + (size_t)i * kw;
    float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
    float4 wa0 = *reinterpret_cast<const float4*>(r + xa);
    float4 wb0 = *reinterpret_cast<const float4*>(r + xb);
    for (int j0 = 0; j0 < kw; j0 += 4) {
      float4 wa1 = *reinterpret_cast<const float4*>(r + xa + j0 + 4);
      float4 wb1 = *reinterpret_cast<const float4*>(r + xb + j0 + 4);
      float4 kk = *reinterpret_cast<const float4*>(kr + j0);
      float wina[8] = {wa0.x, wa0.y, wa0.z, wa0.w, wa1.x, wa1.y, wa1.z, wa1.w};
      float winb[8] = {wb0.x, wb0.y, wb0.z, wb0.w, wb1.x, wb1.y, wb1.z, wb1.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        a[k] = fmaf(kk.x, wina[k], a[k]);
        a[k] = fmaf(kk.y, wina[k + 1], a[k]);
        a[k] = fmaf(kk.z, wina[k + 2], a[k]);
        a[k] = fmaf(kk.w, wina[k + 3], a[k]);
        b[k] = fmaf(kk.x, winb[k], b[k]);
        b[k] = fmaf(kk.y, winb[k + 1], b[k]);
        b[k] = fmaf(kk.z, winb[k + 2], b[k]);
        b[k] = fmaf(kk.w, winb[k + 3], b[k]);
      }
      wa0 = wa1;
      wb0 = wb1;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      acc_a[k] += (double)a[k];
      acc_b[k] += (double)b[k];
    }
  }
  const int y = ty0 + ly;
  if (y < cy) {
    float* o = out + bz * plane_stride + (size_t)y * pitch;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      int xpa = tx0 + xa + k, xpb = tx0 + xb + k;
      if (xpa < pitch) o[xpa] = (xpa >= xpad && xpa < xpad + cx) ? (float)acc_a[k] : 0.f;
      if (xpb < pitch) o[xpb] = (xpb >= xpad && xpb < xpad + cx) ? (float)acc_b[k] : 0.f;
    }
  }
}
__global__ void __launch_bounds__(kConvThreads)
psf_conv_kernel(const float* __restrict__ image, int cy, int cx, int pitch, int xpad,
                const float* __restrict__ kflip, const ConvPlane* __restrict__ planes,
                float* __restrict__ out, size_t plane_stride, int row0, int row1) {
  psf_conv_body(image, cy, cx, pitch, xpad, kflip, planes, out, plane_stride, row0, row1, blockIdx.x, blockIdx.y,
                blockIdx.z);
}

// ----------------------------------------------------------------------------------------------
// K1 + K2 of every source an evaluation moved, as three launches instead of three per source.  An evaluation that
// moves S sources (a position fit's gradient, the first evaluation of a fit) has S independent chains of small
// kernels; launched one by one they are bound by launch latency and fill a fraction of the GPU each.  Here the
// blocks of all chains form one grid per stage: block -> job by binary search in the prefix sums of the jobs'
// block counts (a few hundred entries at most), the job record is copied to shared memory, and the block then
// runs the body of the per-source kernel unchanged -- same arithmetic, same bits.
// ----------------------------------------------------------------------------------------------
struct SpatialBatchJob {
  SpatialJob j;
  // PSF stencil of this job (n_planes == 0: none); image = j.out
  const float* kflip;
  const ConvPlane* planes;
  float* conv_out;
  int n_planes;
  int conv_gx, conv_gy;  // stencil tiles per plane
  int row0, row1;        // rows of the image that can hold a non-zero
  int fill_gx;           // blocks per row of the fill stage
  int pad_[2];
};

// first[k] = first block of job k in this stage's grid, first[n] = grid size
__device__ __forceinline__ int batch_find_job(const int* __restrict__ first, int n, int block) {
  int lo = 0, hi = n - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (__ldg(first + mid) <= block)
      lo = mid;
    else
      hi = mid - 1;
  }
  return lo;
}

template <typename T>
__device__ __forceinline__ void batch_load_job(T* dst, const T* __restrict__ src, int* job_s, const int* first, int n) {
  static_assert(sizeof(T) % sizeof(int) == 0, "job records are copied as ints");
  if (threadIdx.x == 0) *job_s = batch_find_job(first, n, blockIdx.x);
  __syncthreads();
  const int* q = reinterpret_cast<const int*>(src + *job_s);
  int* d = reinterpret_cast<int*>(dst);
  for (int i = threadIdx.x; i < (int)(sizeof(T) / sizeof(int)); i += blockDim.x) d[i] = __ldg(q + i);
  __syncthreads();
}

__global__ void spatial_fill_batch_kernel(const SpatialBatchJob* __restrict__ jobs, const int* __restrict__ first, int n) {
  __shared__ SpatialBatchJob js;
  __shared__ int job_s;
  batch_load_job(&js, jobs, &job_s, first, n);
  const int local = blockIdx.x - __ldg(first + job_s);
  spatial_fill_body(js.j, local % js.fill_gx, local / js.fill_gx);
}

__global__ void spatial_oversample_batch_kernel(const SpatialBatchJob* __restrict__ jobs, const int* __restrict__ first,
                                                int n) {
  __shared__ SpatialBatchJob js;
  __shared__ int job_s;
  batch_load_job(&js, jobs, &job_s, first, n);
  spatial_oversample_body(js.j, blockIdx.x - __ldg(first + job_s));
}

__global__ void __launch_bounds__(kConvThreads)
psf_conv_batch_kernel(const SpatialBatchJob* __restrict__ jobs, const int* __restrict__ first, int n) {
  __shared__ SpatialBatchJob js;
  __shared__ int job_s;
  batch_load_job(&js, jobs, &job_s, first, n);
  const int local = blockIdx.x - __ldg(first + job_s);
  const int per_plane = js.conv_gx * js.conv_gy;
  const int bz = local / per_plane, rem = local - bz * per_plane;
  const SpatialJob& j = js.j;
  psf_conv_body(j.out, j.ecy, j.ecx, j.pitch, j.xpad, js.kflip, js.planes, js.conv_out, (size_t)j.ecy * j.pitch, js.row0,
                js.row1, rem % js.conv_gx, rem / js.conv_gx, bz);
}

// ----------------------------------------------------------------------------------------------
// K0: SpectralModel.integral over the true-energy bins (modeling/models/spectral.py:353-371)
// ----------------------------------------------------------------------------------------------
constexpr int kFluxJobBackground = 3;  // FluxJob::type of a background-factor job (see spectral_flux_kernel)

struct FluxJob {
  int type;      // gpy_spectral_type, or kFluxJobBackground
  int n_bins;
  int n_nodes;
  int sanitize;         // 1: zero the vector when any bin is non-finite (likelihood path), 0: raw integrals
  int pidx[5];          // columns of p[] in the caller's parameter vector (steady-state replay reads them from there)
  int pad_;
  double p[5];          // parameter values in the reference class order
  const double* edges;  // [n_bins + 1]
  const double* nodes;  // [n_bins, n_nodes]
  double* out;          // [n_bins]
};

__device__ __forceinline__ bool isclose0(double val) { return fabs(val) <= 1e-8; }  // np.isclose(val, 0)

// PowerLawSpectralModel.evaluate_integral (spectral.py:1039-1067)
__device__ __forceinline__ double pl_integral(double emin, double emax, double index, double amplitude,
                                              double reference) {
  double val = -1.0 * index + 1.0;
  if (isclose0(val)) return amplitude * reference * log(emax / emin);
  double prefactor = amplitude * reference / val;
  double upper = pow(emax / reference, val);
  double lower = pow(emin / reference, val);
  return prefactor * (upper - lower);
}

__device__ __forceinline__ double spectral_eval(int type, const double* p, double e) {
  if (type == 0) return p[1] * pow(e / p[2], -p[0]);  // PowerLaw.evaluate (spectral.py:1034-1037)
  if (type == 1) {  // LogParabola.evaluate (spectral.py:1903-1908): amplitude, reference, alpha, beta
    double xx = e / p[1];
    double exponent = -p[2] - p[3] * log(xx);
    return p[0] * pow(xx, exponent);
  }
  // ExpCutoffPowerLaw.evaluate (spectral.py:1561-1567): index, amplitude, reference, lambda_, alpha
  double pwl = p[1] * pow(e / p[2], -p[0]);
  double cutoff = exp(-pow(e * p[3], p[4]));
  return pwl * cutoff;
}

__device__ __forceinline__ double log_scale(double v) {
  // LogScale._scale: np.clip(values, tiny, inf) then log (utils/interpolation.py:207-210); NaN stays NaN
  const double tiny = 2.2250738585072014e-308;
  if (v < tiny) v = tiny;
  return log(v);
}

// One block per job.  The trapezoid segments of all bins are independent: one thread per (bin, segment) writes
// its integral to shared memory, then one thread per bin adds its segments in node order (bit-identical to the
// serial per-bin loop below, which remains the fallback when the segments do not fit).  seg_cap = doubles of
// dynamic shared memory.
// params: null, or the caller's parameter vector on the device: the job's values are then read from it (pidx), so a
// replayed launch (CUDA graph of the steady-state step, see evaluate_replay) only needs the new vector.
// reset: null, or the fold kernel's work counter, zeroed here for the launch that follows (saves the memset node of the
// replayed step).
__global__ void spectral_flux_kernel(const FluxJob* __restrict__ jobs, int seg_cap, const double* __restrict__ params,
                                     unsigned* __restrict__ reset) {
  extern __shared__ double seg[];
  if (reset && blockIdx.x == 0 && threadIdx.x == 0) *reset = 0u;
  FluxJob job = jobs[blockIdx.x];
  if (params) {
#pragma unroll
    for (int k = 0; k < 5; ++k)
      if (job.pidx[k] >= 0) job.p[k] = params[job.pidx[k]];
  }
  if (job.type == kFluxJobBackground) {
    // background factors norm * (E_c / E_0)^-tilt per reco bin (PowerLawNormSpectralModel.evaluate,
    // spectral.py:1159-1162; FoVBackgroundModel.evaluate_geom, cube.py:737-756): p = {norm, tilt, reference},
    // edges = reco bin centres [n_nodes], out padded with 1 to n_bins; sanitize == 0: no background model -> 1
    for (int r = threadIdx.x; r < job.n_bins; r += blockDim.x)
      job.out[r] = (job.sanitize && r < job.n_nodes) ? job.p[0] * pow(job.edges[r] / job.p[2], -job.p[1]) : 1.0;
    return;
  }
  __shared__ int bad;
  if (threadIdx.x == 0) bad = 0;
  const int nseg = job.n_nodes - 1;
  // staged: the model is evaluated once per node (one thread each), then one thread per segment integrates between
  // two stored values; the buffer holds the node values followed by the segment integrals
  const bool staged = job.type != 0 && job.n_bins * (nseg + job.n_nodes) <= seg_cap;
  double* yv = seg + job.n_bins * nseg;  // [n_bins][n_nodes]
  if (staged) {
    for (int w = threadIdx.x; w < job.n_bins * job.n_nodes; w += blockDim.x) yv[w] = spectral_eval(job.type, job.p, job.nodes[w]);
    __syncthreads();
    for (int w = threadIdx.x; w < job.n_bins * nseg; w += blockDim.x) {
      const int t = w / nseg, k = w - t * nseg;
      const double* nd = job.nodes + (size_t)t * job.n_nodes + k;
      const double e0 = nd[0], e1 = nd[1];
      const double y0 = yv[t * job.n_nodes + k], y1 = yv[t * job.n_nodes + k + 1];
      double index = -log_scale(y0 / y1) / log_scale(e0 / e1);
      if (isnan(index)) index = INFINITY;
      seg[w] = pl_integral(e0, e1, index, y0, e0);
    }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < job.n_bins; t += blockDim.x) {
    double emin = job.edges[t], emax = job.edges[t + 1];
    double result;
    if (job.type == 0) {
      result = pl_integral(emin, emax, job.p[0], job.p[1], job.p[2]);
    } else if (staged) {
      result = 0.0;
      for (int k = 0; k < nseg; ++k) result += seg[t * nseg + k];
    } else {
      // integrate_spectrum + trapz_loglog (spectral.py:103-164, utils/integrate.py:8-49)
      const double* nd = job.nodes + (size_t)t * job.n_nodes;
      double e0 = nd[0], y0 = spectral_eval(job.type, job.p, e0);
      result = 0.0;
      for (int k = 1; k < job.n_nodes; ++k) {
        double e1 = nd[k], y1 = spectral_eval(job.type, job.p, e1);
        double index = -log_scale(y0 / y1) / log_scale(e0 / e1);
        if (isnan(index)) index = INFINITY;
        result += pl_integral(e0, e1, index, y0, e0);
        e0 = e1;
        y0 = y1;
      }
    }
    if (!isfinite(result)) bad = 1;
    job.out[t] = result;
  }
  if (!job.sanitize) return;
  // A non-finite flux in ANY true-energy bin makes the source's whole npred non-finite after the edisp
  // matmul (0 * NaN = NaN), and WcsNDMap.stack(nan_to_num=True) then drops the source: zero flux here.
  __syncthreads();
  if (bad)
    for (int t = threadIdx.x; t < job.n_bins; t += blockDim.x) job.out[t] = 0.0;
}

// ----------------------------------------------------------------------------------------------
// K3 + K4 fused: apply_exposure -> apply_edisp -> stack -> + background -> clip -> masked Cash
//   (datasets/evaluator.py:346-377, datasets/utils.py:66-84, maps/wcs/ndmap.py:1138-1181,
//    datasets/map.py:775-822, stats/fit_statistics.py:281-293, fit_statistics_cython.pyx:55-85)
// ----------------------------------------------------------------------------------------------
struct DatasetDev {
  int nx, ny, nT, nR;
  int nRp;        // nR rounded up to a multiple of 8
  int n_groups;   // edisp matrices: 0 = dataset edisp, 1 = diagonal response
  int tile_begin, n_tiles;
  long long P;    // ny * nx
  const float* exposure;    // [nT, P]
  const float* background;  // [nR, P] or null
  const float* counts;      // [nR, P] or null
  const uint8_t* mask;      // [nR, P] or null
  const float* weight;      // [nR, P] or null: float mask of the weighted Cash (generic kernel only)
  const double* edisp;      // [n_groups, nT, nRp] zero padded
  const int* band;          // [n_groups, nRp / 8, 4]: first / last+1 true bin with non-zero pdf per reco chunk, row offset
                            // of the chunk in the band-compressed matrix, pad
  const double* pdfc;       // band-compressed edisp: for every (group, chunk) the rows [lo, hi) x 8 reco bins, contiguous
  int pdfc_rows;            // total rows (of 8 doubles) in pdfc
  int compact;              // 1: counts are uint16 [nR, P], mask is bit-packed (bit p & 7 of byte p >> 3, little endian)
                            //    with rows of mask_row_bytes (a multiple of 16, zero padded) -- the C5 storage format
  int mask_row_bytes, pad3_;
  const double* ecenter;    // [nR]
  const void* tmaps;        // 4 CUtensorMap (exposure, background, counts, mask; 128 B each) or null: 1-D bulk copies
};

struct InstDev {  // one source in one parameter vector
  int x0, y0, cx, cy;  // evaluator cutout in parent pixels
  int x0a, pitch;      // slot column origin (x0 rounded down to a multiple of 4) and row pitch
  int planes;          // nT, or 1 when the PSF is not applied (broadcast over energy)
  int group;           // edisp group
  int flux_off;        // offset of F[nT] in the flux buffer
  int plane_stride;    // cy * pitch, or 0 when planes == 1 (same image for every energy)
  const float* conv;   // [planes, cy, pitch]
};

struct EvalDev {  // one (dataset, parameter vector)
  int inst_begin, inst_count;
  int bkg_enabled;
  int differs_all;             // batched calls: every tile differs from vector 0's (background parameters changed)
  int diff_begin, diff_count;  // batched calls: indices (into insts) of the sources that differ from vector 0's
  double bkg_norm, bkg_tilt, bkg_ref;
};

struct FoldParams {
  const DatasetDev* ds;
  const int* tile_ds;       // dataset index of every tile
  const EvalDev* evals;     // [n_datasets * B], index d * B + b
  const InstDev* insts;
  const double* flux;
  double* partials;         // [B, n_tiles]
  double* npred_out;        // [nR, P] or null (single dataset, B == 1)
  int B, n_tiles, n_datasets;
  int include_background;   // add background model and clip at 0 (MapDataset.npred) or signal only
  int max_inst;             // capacity of the per-block overlap list
  int smem_nT, smem_nR, smem_nRp, smem_G;  // largest axes of the launch: fixed shared-memory layout (bulk-copy kernel)
  int smem_pdfc_rows, pad3_;
  const unsigned char* desc;  // [work items][desc_layout(smem_nT, smem_nRp).bytes] item descriptors (prep_items_kernel)
  // shared-memory byte offsets (fold_tma_layout) and descriptor offsets (desc_layout), computed once on the host:
  // as kernel parameters they are constant-bank operands instead of arithmetic the compiler rematerialises at
  // every shared-memory access (13 % of the kernel's instructions before)
  unsigned lay_es, lay_bs, lay_cs, lay_ms, lay_vs, lay_pdf, lay_desc, lay_dyn, lay_bars;
  int dl_inst, dl_over, dl_bytes;
  const double* bkgfac;     // [n_datasets * B][smem_nRp] background factors norm (E/E0)^-tilt (K0), index d * B + b
  unsigned* counter;        // work-stealing counter (0 at launch): items beyond a CTA's first are claimed in order; null: static
  const int* work;          // [*n_work] item indices to evaluate, or null: all B * n_tiles items (B == 1)
  const int* n_work;
  const float* zeros;       // >= 16 bytes of zeros: load target of (thread, source) pairs outside the cutout
  int pad_;
  double log_trunc;         // log((double)(float)1e-25): Cython narrows the constant to C float
  const double2* logtab;    // fast_log table, 128 entries
  const double2* logtab64;  // 64-entry variant (fold_cash_t64_kernel)
};

constexpr int kFoldTPX = 128;      // pixels per tile
constexpr int kFoldThreads = 256;
constexpr int kFoldTCH = 16;       // true-energy bins staged in shared memory per pass
#ifndef GPY_FOLD_MINBLOCKS
#define GPY_FOLD_MINBLOCKS 3
#endif

// Shared memory of one block: v chunk [G][TCH][TPX] f64, pdf [G][nT][nRp] f64, bkgfac [nRp], scratch [8],
// overlap list [max_inst] InstDev.
// (one edisp matrix resident at a time, whatever the number of matrices of the dataset: see the kernel)
__host__ __device__ inline size_t fold_smem_bytes(int nT, int nRp, int max_inst) {
  return ((size_t)kFoldTCH * kFoldTPX + (size_t)nT * nRp + nRp + 8) * sizeof(double) +
         (size_t)max_inst * (sizeof(InstDev) + sizeof(int));
}

__device__ __forceinline__ double block_sum_256(double v, double* scratch) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  double r = 0.0;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 0; w < kFoldThreads / 32; ++w) r += scratch[w];
  }
  return r;
}

// One block = one tile of 128 consecutive pixels x one parameter vector.
//   phase 1 (per pass of 16 true-energy bins): thread (pixel quad q, energy slot) computes
//       v[g][t][p] = 1e4 * exposure[t,p] * sum_s F_s[t] * conv_s[t,p]          (fp64, float4 loads)
//     for the sources whose cutout overlaps the tile, into shared memory;
//   phase 2: thread (pixel pair, reco chunk of 8) accumulates acc[r] += pdf[t][r] * v[t][p] over the
//     non-zero band of the edisp matrix; accumulators stay in registers across the passes;
//   epilogue: + background * norm (E/E0)^-tilt, clip, masked Cash term, block reduction.
// Edisp matrices ("groups": with position-dependent IRFs every evaluator has the matrix of its nearest IRF pixel,
// irf/edisp/map.py:369-401, so a dataset can hold as many matrices as it has sources): the tile works through the
// distinct matrices of ITS overlapping sources one after the other -- matrix into shared memory, phase 1 and phase 2 of
// the sources that use it, accumulators kept -- so shared memory does not grow with the number of matrices, and a
// tile no source reaches goes straight to the epilogue.
__global__ void __launch_bounds__(kFoldThreads, GPY_FOLD_MINBLOCKS) fold_cash_generic_kernel(const FoldParams P) {
  extern __shared__ __align__(16) double fold_smem[];
  const int b = blockIdx.x % P.B;
  const int tile = blockIdx.x / P.B;
  const int d = P.tile_ds[tile];
  const DatasetDev D = P.ds[d];
  const EvalDev ev = P.evals[(size_t)d * P.B + b];
  const int nT = D.nT, nR = D.nR, nRp = D.nRp;
  const long long p0 = (long long)(tile - D.tile_begin) * kFoldTPX;
  const int npx = (int)min((long long)kFoldTPX, D.P - p0);
  const int tid = threadIdx.x;

  double* vs = fold_smem;                                      // [TCH][TPX]
  double* pdf_s = vs + (size_t)kFoldTCH * kFoldTPX;            // [nT][nRp]: the matrix being worked on
  double* bkgfac = pdf_s + (size_t)nT * nRp;                   // [nRp]
  double* scratch = bkgfac + nRp;                              // [8]
  InstDev* slist = reinterpret_cast<InstDev*>(scratch + 8);    // [max_inst]
  int* glist = reinterpret_cast<int*>(slist + P.max_inst);     // [max_inst]: distinct matrices of slist, ascending
  __shared__ int n_list, n_glist;

  if (tid < nRp) {
    double fac = 1.0;
    if (ev.bkg_enabled && tid < nR)  // PowerLawNormSpectralModel.evaluate (spectral.py:1159-1162)
      fac = ev.bkg_norm * pow(D.ecenter[tid] / ev.bkg_ref, -ev.bkg_tilt);
    bkgfac[tid] = fac;
  }
  // sources whose cutout overlaps the rows (and, for a single-row tile, the columns) of this tile
  const int y_first = (int)(p0 / D.nx), y_last = (int)((p0 + npx - 1) / D.nx);
  const int x_first = (int)(p0 - (long long)y_first * D.nx);
  const int x_last = (int)(p0 + npx - 1 - (long long)y_last * D.nx);
  if (tid < 32) {
    int n = 0;
    for (int base = 0; base < ev.inst_count; base += 32) {
      const int i = base + tid;
      bool ov = false;
      InstDev in;
      if (i < ev.inst_count) {
        in = P.insts[ev.inst_begin + i];
        ov = in.y0 <= y_last && in.y0 + in.cy > y_first;
        if (ov && y_first == y_last) ov = in.x0 <= x_last && in.x0 + in.cx > x_first;
      }
      const unsigned m = __ballot_sync(0xffffffffu, ov);
      if (ov) slist[n + __popc(m & ((1u << tid) - 1u))] = in;
      n += __popc(m);
    }
    if (tid == 0) {
      n_list = n;
      int ng = 0;
      for (int l = 0; l < n; ++l) {  // sorted insertion of the distinct group indices (a handful per tile)
        const int g = slist[l].group;
        int k = 0;
        while (k < ng && glist[k] < g) ++k;
        if (k < ng && glist[k] == g) continue;
        for (int m = ng; m > k; --m) glist[m] = glist[m - 1];
        glist[k] = g;
        ++ng;
      }
      n_glist = ng;
    }
  }
  __syncthreads();
  const int nl = n_list, ngl = n_glist;
  int loaded = -1;  // matrix currently in pdf_s

  // phase-1 coordinates: 32 pixel quads x 8 energy slots
  const int q = tid & 31, slot = tid >> 5;
  const long long pq = p0 + 4 * q;
  const bool vec4 = (D.nx % 4) == 0;      // quads are row-aligned and 16-byte aligned
  int qy[4], qx[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const long long p = pq + k;
    const bool live = 4 * q + k < npx;
    qy[k] = live ? (int)(p / D.nx) : -1;
    qx[k] = live ? (int)(p - (long long)qy[k] * D.nx) : 0;
  }
  // phase-2 coordinates: 64 pixel pairs x 4 reco-chunk slots
  const int pp = tid & 63, cslot = tid >> 6;
  const int nchunks = nRp >> 3;

  double stat = 0.0;
  for (int cbase = 0; cbase < nchunks; cbase += 4) {  // one pass for nR <= 32
    const int c = cbase + cslot;
    double acc0[8], acc1[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc0[j] = acc1[j] = 0.0;

    for (int lg = 0; lg < ngl; ++lg) {
    const int g = glist[lg];
    if (g != loaded) {  // (block uniform; the previous user of pdf_s finished behind the barrier that ends its pass)
      const double* src = D.edisp + (size_t)g * nT * nRp;
      for (int i = tid; i < nT * nRp; i += kFoldThreads) pdf_s[i] = __ldg(src + i);
      loaded = g;
      // (visible to phase 2 through the barrier after phase 1)
    }
    for (int t0 = 0; t0 < nT; t0 += kFoldTCH) {
      // ---- phase 1 ----
      {
#pragma unroll
        for (int tt = 0; tt < kFoldTCH / 8; ++tt) {
          const int tl = slot + 8 * tt, t = t0 + tl;
          double s[4] = {0.0, 0.0, 0.0, 0.0};
          if (t < nT && qy[0] >= 0) {
            for (int l = 0; l < nl; ++l) {
              const InstDev& in = slist[l];
              if (in.group != g) continue;
              const int tp = in.planes == 1 ? 0 : t;
              const double F = P.flux[in.flux_off + t];
              const float* plane = in.conv + (size_t)tp * in.cy * in.pitch;
              const int ly = qy[0] - in.y0, lx = qx[0] - in.x0a;
              if (vec4) {
                if (ly < 0 || ly >= in.cy || lx < 0 || lx >= in.pitch) continue;
                const float4 cv = __ldg(reinterpret_cast<const float4*>(plane + (size_t)ly * in.pitch + lx));
                const double a0 = F * (double)cv.x, a1 = F * (double)cv.y, a2 = F * (double)cv.z,
                             a3 = F * (double)cv.w;
                // WcsNDMap.stack(nan_to_num=True): non-finite source values do not enter the sum
                s[0] += isfinite(a0) ? a0 : 0.0;
                s[1] += isfinite(a1) ? a1 : 0.0;
                s[2] += isfinite(a2) ? a2 : 0.0;
                s[3] += isfinite(a3) ? a3 : 0.0;
              } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  if (qy[k] < 0) continue;
                  const int lyk = qy[k] - in.y0, lxk = qx[k] - in.x0;
                  if (lyk < 0 || lyk >= in.cy || lxk < 0 || lxk >= in.cx) continue;
                  const double a = F * (double)__ldg(plane + (size_t)lyk * in.pitch + (qx[k] - in.x0a));
                  s[k] += isfinite(a) ? a : 0.0;
                }
              }
            }
            const float* erow = D.exposure + (size_t)t * D.P + pq;
            if (vec4) {
              const float4 e = __ldg(reinterpret_cast<const float4*>(erow));
              s[0] = (s[0] * (double)e.x) * 1e4;
              s[1] = (s[1] * (double)e.y) * 1e4;
              s[2] = (s[2] * (double)e.z) * 1e4;
              s[3] = (s[3] * (double)e.w) * 1e4;
            } else {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                if (qy[k] >= 0) s[k] = (s[k] * (double)__ldg(erow + k)) * 1e4;
            }
          }
          double* dst = vs + (size_t)tl * kFoldTPX + 4 * q;
          *reinterpret_cast<double2*>(dst) = make_double2(s[0], s[1]);
          *reinterpret_cast<double2*>(dst + 2) = make_double2(s[2], s[3]);
        }
      }
      __syncthreads();
      // ---- phase 2 ----
      if (c < nchunks) {
        {
          const int blo = D.band[(g * nchunks + c) * 4], bhi = D.band[(g * nchunks + c) * 4 + 1];
          const int tlo = max(blo, t0), thi = min(bhi, t0 + kFoldTCH);
          const double* vrow = vs + 2 * pp;
          const double* prow = pdf_s + 8 * c;
          for (int t = tlo; t < thi; ++t) {
            const double2 v = *reinterpret_cast<const double2*>(vrow + (size_t)(t - t0) * kFoldTPX);
            const double2 m0 = *reinterpret_cast<const double2*>(prow + (size_t)t * nRp);
            const double2 m1 = *reinterpret_cast<const double2*>(prow + (size_t)t * nRp + 2);
            const double2 m2 = *reinterpret_cast<const double2*>(prow + (size_t)t * nRp + 4);
            const double2 m3 = *reinterpret_cast<const double2*>(prow + (size_t)t * nRp + 6);
            const double m[8] = {m0.x, m0.y, m1.x, m1.y, m2.x, m2.y, m3.x, m3.y};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              acc0[j] = fma(v.x, m[j], acc0[j]);
              acc1[j] = fma(v.y, m[j], acc1[j]);
            }
          }
        }
      }
      __syncthreads();
    }
    }

    // ---- epilogue for reco chunk c: background, clip, Cash ----
    if (c < nchunks) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int r = 8 * c + j;
        if (r >= nR) break;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const int lp = 2 * pp + k;
          if (lp >= npx) continue;
          const size_t idx = (size_t)r * D.P + p0 + lp;
          double mu = k == 0 ? acc0[j] : acc1[j];
          if (P.include_background) {
            if (D.background) mu += (double)__ldg(D.background + idx) * bkgfac[r];
            if (mu < 0.0) mu = 0.0;  // npred_total.data[npred_total.data < 0.0] = 0 (map.py:787)
          }
          if (P.npred_out) P.npred_out[idx] = mu;
          if (D.counts) {
            bool m = true;
            if (D.mask)
              m = D.compact ? ((__ldg(D.mask + (size_t)r * D.mask_row_bytes + ((p0 + lp) >> 3)) >> ((p0 + lp) & 7)) & 1) != 0
                            : __ldg(D.mask + idx) != 0;
            if (m) {
              const double n = D.compact ? (double)__ldg(reinterpret_cast<const unsigned short*>(D.counts) + idx)
                                         : (double)__ldg(D.counts + idx);
              double npr, lognpr = 0.0;
              if (mu > kTrunc) {
                npr = mu;
                if (n > 0.0) lognpr = log(mu);
              } else {
                npr = kTrunc;
                lognpr = P.log_trunc;
              }
              if (D.weight) {  // weighted_cash_sum_cython (fit_statistics_cython.pyx:17-51)
                const double w = (double)__ldg(D.weight + idx);
                if (w > 0.0) {
                  stat += w * npr;
                  if (n > 0.0) stat -= w * n * lognpr;
                }
              } else {
                stat += npr;
                if (n > 0.0) stat -= n * lognpr;
              }
            }
          }
        }
      }
    }
  }
  double total = block_sum_256(stat, scratch);
  if (tid == 0) P.partials[(size_t)b * P.n_tiles + tile] = total;
}

// ----------------------------------------------------------------------------------------------
// Bulk-copy (TMA engine) staged, persistent variant of the fused fold + Cash kernel.
//
// Each CTA loops over work items (tile of 128 pixels, parameter vector).  The exposure tile [nT][128],
// and the background / counts / mask tiles [nR][128] of the NEXT item are fetched by 1-D
// cp.async.bulk copies (one 512-byte row segment each, completion on an mbarrier) while the current
// item is being computed, so phase 1, phase 2 and the Cash epilogue read shared memory only; the
// per-source PSF-convolved cubes (irregular cutouts) are read with 128-bit __ldg loads.
// Requirements (checked on the host, otherwise the generic kernel runs): nx % 4 == 0, P % 16 == 0,
// 16-byte aligned cubes.
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ float4 ldg_f4_hint(const float* p, uint64_t pol) {
  float4 v;
  asm("ld.global.nc.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ void prefetch_l2_keep(const void* p) {
  asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(p));
}
__device__ __forceinline__ void bulk_g2s_hint(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
      : "memory");
}

__device__ __forceinline__ void tma_load_2d(void* dst, const void* tmap, int c0, int c1, uint64_t* bar, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(
          smem_u32(dst)),
      "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "l"(pol)
      : "memory");
}

// ---- batched calls: evaluate every distinct item once ----------------------------------------------------------
// In a profile scan or a finite-difference gradient most (tile, vector) items are identical to the item of vector 0
// on the same tile: the varied source does not reach the tile.  The host lists, per vector, the sources that differ
// from vector 0's (EvalDev::diff_*: another PSF-convolved cube, flux vector or edisp group, or present in one of the
// two only); an item is a duplicate when none of them overlaps its tile and the background parameters are equal.
// The kernels below flag the duplicates and build, in a deterministic order, the list of items to evaluate;
// reduce_stat_kernel reads the partial of (tile, 0) for the duplicates.
__global__ void __launch_bounds__(256)
dedup_flag_kernel(const DatasetDev* __restrict__ ds, const int* __restrict__ tile_ds, const EvalDev* __restrict__ evals,
                  const InstDev* __restrict__ insts, const int* __restrict__ diffs, int B, int n_tiles,
                  unsigned char* __restrict__ dup, int* __restrict__ count) {
  const int lane = threadIdx.x & 31;
  const int tile = (int)(((long long)blockIdx.x * 256 + threadIdx.x) >> 5);  // one warp per tile
  if (tile >= n_tiles) return;
  const int d = tile_ds[tile];
  const DatasetDev& D = ds[d];
  const long long pix0 = (long long)(tile - D.tile_begin) * kFoldTPX;
  const int npx = (int)min((long long)kFoldTPX, D.P - pix0);
  const int y_first = (int)(pix0 / D.nx), y_last = (int)((pix0 + npx - 1) / D.nx);
  const int x_first = (int)(pix0 % D.nx), x_last = (int)((pix0 + npx - 1) % D.nx);
  int n = 0;
  for (int b0 = 0; b0 < B; b0 += 32) {
    const int b = b0 + lane;
    bool unique = false;
    if (b < B) {
      const EvalDev& ev = evals[(size_t)d * B + b];
      unique = b == 0 || ev.differs_all;
      for (int k = 0; k < ev.diff_count && !unique; ++k) {
        const InstDev& in = insts[diffs[ev.diff_begin + k]];
        bool ov = in.y0 <= y_last && in.y0 + in.cy > y_first;
        if (ov && y_first == y_last) ov = in.x0 <= x_last && in.x0 + in.cx > x_first;
        unique = ov;
      }
      dup[(size_t)tile * B + b] = unique ? 0 : 1;
    }
    n += __popc(__ballot_sync(0xffffffffu, unique));
  }
  if (lane == 0) count[tile] = n;
}

__global__ void __launch_bounds__(1024) dedup_scan_kernel(const int* __restrict__ count, int n_tiles, int* __restrict__ offset,
                                                          int* __restrict__ n_work) {
  __shared__ int sm[1024];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < n_tiles; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < n_tiles ? count[i] : 0;
    sm[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {  // Hillis-Steele inclusive scan
      const int add = threadIdx.x >= o ? sm[threadIdx.x - o] : 0;
      __syncthreads();
      sm[threadIdx.x] += add;
      __syncthreads();
    }
    if (i < n_tiles) offset[i] = carry + sm[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry += sm[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) *n_work = carry;
}

__global__ void __launch_bounds__(256)
dedup_scatter_kernel(const unsigned char* __restrict__ dup, const int* __restrict__ offset, int B, int n_tiles,
                     int* __restrict__ work) {
  const int lane = threadIdx.x & 31;
  const int tile = (int)(((long long)blockIdx.x * 256 + threadIdx.x) >> 5);  // one warp per tile, vectors in order
  if (tile >= n_tiles) return;
  int pos = offset[tile];
  for (int b0 = 0; b0 < B; b0 += 32) {
    const int b = b0 + lane;
    const bool unique = b < B && dup[(size_t)tile * B + b] == 0;
    const unsigned m = __ballot_sync(0xffffffffu, unique);
    if (unique) work[pos + __popc(m & ((1u << lane) - 1u))] = tile * B + b;
    pos += __popc(m);
  }
}

// ---- per-item work descriptors --------------------------------------------------------------------------
// prep_items_kernel (one warp per (tile, vector) item, after K0) resolves everything an item needs besides the
// data cubes and the per-step numbers: the sources whose cutout overlaps the tile (first kFoldLMax as InstDev copies,
// further ones as indices).  Nothing in a descriptor depends on the spectral or background parameters -- the fold
// kernel stages the flux vectors and background factors K0 wrote (a few KB) in shared memory one item ahead -- so
// the host re-runs this kernel only when a cutout, a slot or the source list changes, not on the steady-state steps
// of a fit.  The fold kernel fetches the descriptor of the NEXT item with one bulk copy while it computes the
// current one, so its prologue is a single mbarrier wait instead of a serial list build over global memory.
constexpr int kFoldLMax = 4;       // overlapping sources held in registers by the fold kernel
constexpr int kDescOverflow = 60;  // further overlapping sources per item (indices; slow path)
struct DescLayout {
  int inst, over, bytes;  // byte offsets inside a descriptor; header = int4 {nl, n_over, d * B + b, 0}
};
__host__ __device__ inline DescLayout desc_layout() {
  DescLayout L;
  L.inst = 16;
  L.over = L.inst + kFoldLMax * (int)sizeof(InstDev);
  L.bytes = (L.over + kDescOverflow * 4 + 15) / 16 * 16;
  return L;
}

struct PrepParams {
  const DatasetDev* ds;
  const int* tile_ds;
  const EvalDev* evals;
  const InstDev* insts;
  unsigned char* desc;  // [number of work items][desc bytes]
  const int* work;          // prep_items_kernel: items to describe ([*n_work]), or null: all of them
  const int* n_work;
  int B, n_tiles;
};

__global__ void __launch_bounds__(256) prep_items_kernel(const PrepParams P) {
  const int lane = threadIdx.x & 31;
  const long long w = ((long long)blockIdx.x * 256 + threadIdx.x) >> 5;  // position in the worklist = descriptor slot
  if (w >= (P.work ? (long long)*P.n_work : (long long)P.B * P.n_tiles)) return;
  const long long item = P.work ? P.work[w] : w;
  const DescLayout L = desc_layout();
  unsigned char* out = P.desc + (size_t)w * L.bytes;
  const int b = (int)(item % P.B);
  const int tile = (int)(item / P.B);
  const int d = P.tile_ds[tile];
  const DatasetDev& D = P.ds[d];
  const EvalDev& ev = P.evals[(size_t)d * P.B + b];
  const int p0 = (tile - D.tile_begin) * kFoldTPX;
  const int npx = min(kFoldTPX, (int)D.P - p0);
  const int y_first = p0 / D.nx, y_last = (p0 + npx - 1) / D.nx;
  const int x_first = p0 - y_first * D.nx, x_last = p0 + npx - 1 - y_last * D.nx;
  InstDev* oinst = reinterpret_cast<InstDev*>(out + L.inst);
  int* oover = reinterpret_cast<int*>(out + L.over);
  int n = 0;
  for (int base = 0; base < ev.inst_count; base += 32) {
    const int i = base + lane;
    bool ov = false;
    InstDev in;
    if (i < ev.inst_count) {
      in = P.insts[ev.inst_begin + i];
      ov = in.y0 <= y_last && in.y0 + in.cy > y_first;
      if (ov && y_first == y_last) ov = in.x0 <= x_last && in.x0 + in.cx > x_first;
    }
    const unsigned m = __ballot_sync(0xffffffffu, ov);
    const int pos = n + __popc(m & ((1u << lane) - 1u));
    if (ov) {
      if (pos < kFoldLMax)
        oinst[pos] = in;
      else if (pos < kFoldLMax + kDescOverflow)
        oover[pos - kFoldLMax] = ev.inst_begin + i;
    }
    n += __popc(m);
  }
  __syncwarp();
  const int nl = min(n, kFoldLMax);
  if (lane == 0)  // header: sources in registers, overflow sources, (dataset, vector) index of the background factors
    *reinterpret_cast<int4*>(out) = make_int4(nl, min(max(n - kFoldLMax, 0), kDescOverflow), d * P.B + b, 0);
}

#ifndef GPY_ABLATE  // profiling only (wrong results): 1 slot-cube loads hit a zero word, 2 no phase 2, 4 no Cash pass
#define GPY_ABLATE 0
#endif
#ifndef GPY_PF_PASS
#define GPY_PF_PASS 0
#endif
// log(x) for normal finite x > 0: x = 2^e m, m in [1, 2); m = m_hi (1 + r) with 1 / m_hi from a table indexed by the
// top 7 mantissa bits, |r| < 2^-8; log(x) = e ln2 + log(m_hi) + log1p(r), log1p by its degree-7 Taylor polynomial
// (|error| < 1e-17 absolute + the table's rounding, ~1e-16).  Straight-line code: calls interleave, unlike libdevice's
// log with its data-dependent branches.  Table: double2 {1 / m_hi, -log(1 / m_hi)} x 128 (built by the host).
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

constexpr int kTmaTCH = 32;  // true-energy bins per pass of the bulk-copy kernel (2 passes at nT = 60)
constexpr int kBandSmem = 128;  // ints of the band table kept in shared memory (larger tables are read from global)
struct FoldSmem {  // byte offsets of the fixed shared-memory layout
  size_t es, bs, cs, ms, vs, pdf, desc, dyn, bars, total;
};
__host__ __device__ inline FoldSmem fold_tma_layout(int nT, int nR, int nRp, int G, int pdfc_rows, bool compact = false) {
  FoldSmem L;  // compact: every dataset of the launch stores uint16 counts and a bit-packed mask
  size_t o = 0;
  L.es = o;    o += (size_t)nT * kFoldTPX * 4;
  L.bs = o;    o += (size_t)nR * kFoldTPX * 4;
  L.cs = o;    o += (size_t)nR * kFoldTPX * (compact ? 2 : 4);
  L.ms = o;    o += compact ? (size_t)nR * 16 : ((size_t)nR * kFoldTPX + 15) / 16 * 16;
  L.vs = o;    o += (size_t)G * kTmaTCH * kFoldTPX * 8;
  L.pdf = o;   o += (size_t)pdfc_rows * 64;
  L.desc = o;  o += (size_t)2 * desc_layout().bytes;  // double buffered item descriptors
  L.dyn = o;   o += (size_t)2 * (kFoldLMax * nT + nRp) * 8;  // double buffered flux rows + background factors
  L.bars = (o + 7) / 8 * 8;
  L.total = L.bars + 4 * 8;
  return L;
}

// COMPACT: every dataset of the launch stores uint16 counts and a bit-packed mask (DatasetDev::compact).  A template
// parameter, not a run-time flag: the block-uniform branch in the Cash pass cost 7 % of the kernel.
template <bool WRITE_NPRED, bool DYNAMIC, bool COMPACT>
__global__ void __launch_bounds__(kFoldThreads, 2) fold_cash_tma_kernel(const FoldParams P) {
  extern __shared__ __align__(128) unsigned char fold_raw[];
  float* es = reinterpret_cast<float*>(fold_raw + P.lay_es);
  float* bs = reinterpret_cast<float*>(fold_raw + P.lay_bs);
  float* cs = reinterpret_cast<float*>(fold_raw + P.lay_cs);
  unsigned char* ms = fold_raw + P.lay_ms;
  double* vs = reinterpret_cast<double*>(fold_raw + P.lay_vs);
  double* pdf_s = reinterpret_cast<double*>(fold_raw + P.lay_pdf);
  unsigned char* desc_s = fold_raw + P.lay_desc;
  double* dyn_s = reinterpret_cast<double*>(fold_raw + P.lay_dyn);  // [2][kFoldLMax * smem_nT + smem_nRp]
  DescLayout DL;
  DL.inst = P.dl_inst;
  DL.over = P.dl_over;
  DL.bytes = P.dl_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(fold_raw + P.lay_bars);  // [0] exposure, [1] epilogue inputs, [2], [3] descriptor buffers
  __shared__ double wsum[kFoldThreads / 32];
  __shared__ int2 nxt_s;         // {worklist position, item} claimed for the iteration after the next
  __shared__ DatasetDev Ds;      // table entry of the dataset the current item belongs to: no global loads per item
  __shared__ __align__(16) int band_s[kBandSmem];  // read as int4
  __shared__ double2 logtab_s[128];
  if (threadIdx.x < 128) logtab_s[threadIdx.x] = P.logtab[threadIdx.x];  // published by the barrier behind the mbarrier initialisation

  const int tid = threadIdx.x;
  const int n_items = P.B * P.n_tiles;
  // L2 policy: the per-source PSF-convolved cubes are re-read by every evaluation (keep), the dataset cubes stream
  const uint64_t pol_keep = l2_policy_evict_last(), pol_stream = l2_policy_evict_first();
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&bars[2], 1);
    mbar_init(&bars[3], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // Geometry of a work item without touching global memory: the table entry of the current dataset lives in shared
  // memory (Ds); only an item of ANOTHER dataset (the CTA crosses a dataset boundary) reads the table in global memory.
  auto geom = [&](int item, int& p0, int& npx) -> const DatasetDev* {
    const int tile = DYNAMIC ? item : item / P.B;  // dynamic launches have B == 1
    const DatasetDev* Dn = &Ds;
    if (tile < Ds.tile_begin || tile >= Ds.tile_begin + Ds.n_tiles) Dn = P.ds + P.tile_ds[tile];
    p0 = (tile - Dn->tile_begin) * kFoldTPX;
    npx = min(kFoldTPX, (int)Dn->P - p0);
    return Dn;
  };
  // producers: thread 0 arms the mbarrier with the byte count BEFORE a CTA barrier, then every warp issues its
  // share of the 1-D bulk copies (one 512-byte row segment per lane), so the ~70-cycle issue cost of a copy is
  // spread over the 8 warps instead of sitting on warp 0's critical path.  With tensor maps thread 0 arms and issues
  // one 2-D tensor copy per cube in issue_*.
  // Producer duties sit on lane 0 of three different warps, so their issue costs (an expect_tx + tensor copy is a
  // few hundred cycles) run side by side instead of queueing on one thread that the other 255 then wait for.
  constexpr int kEpiTid = 32, kClaimTid = 64, kExpoTid = 96;
  auto arm_expo = [&](int item) {
    if (tid != kExpoTid) return;
    int p0, npx;
    const DatasetDev* D = geom(item, p0, npx);
    if (!D->tmaps) mbar_arrive_expect_tx(&bars[0], (uint32_t)npx * 4u * (uint32_t)D->nT);
  };
  auto issue_expo = [&](int item) {
    int p0, npx;
    const DatasetDev* D = geom(item, p0, npx);
    const uint32_t row = (uint32_t)npx * 4u;
    if (D->tmaps) {  // one 2-D tensor copy of the whole [nT][128] box (out-of-range columns are zero filled)
      if (tid == kExpoTid) {
        mbar_arrive_expect_tx(&bars[0], (uint32_t)kFoldTPX * 4u * (uint32_t)D->nT);
        tma_load_2d(es, D->tmaps, p0, 0, &bars[0], pol_stream);
      }
      return;
    }
    const int nT = D->nT;
    const float* expo = D->exposure;
    const size_t PP = (size_t)D->P;
    for (int t = (tid & 31) * (kFoldThreads / 32) + (tid >> 5); t < nT; t += kFoldThreads)  // row handled by this lane
      bulk_g2s_hint(es + t * kFoldTPX, expo + (size_t)t * PP + p0, row, &bars[0], pol_stream);
  };
  auto arm_epi = [&](int item) {
    if (tid != kEpiTid) return;
    int p0, npx;
    const DatasetDev* D = geom(item, p0, npx);
    const uint32_t row = (uint32_t)npx * 4u;
    uint32_t total = 0;
    if (D->tmaps) return;  // armed together with the tensor copies, after the barrier
    if (D->background && P.include_background) total += row * (uint32_t)D->nR;
    if (D->counts) total += (COMPACT ? row / 2u : row) * (uint32_t)D->nR;
    if (D->counts && D->mask) total += (COMPACT ? 16u : (uint32_t)npx) * (uint32_t)D->nR;
    mbar_arrive_expect_tx(&bars[1], total);
  };
  auto issue_epi = [&](int item) {
    int p0, npx;
    const DatasetDev* D = geom(item, p0, npx);
    const uint32_t row = (uint32_t)npx * 4u;
    const bool has_bkg = D->background && P.include_background;
    if (D->tmaps) {
      if (tid == kEpiTid) {
        const unsigned char* tm = reinterpret_cast<const unsigned char*>(D->tmaps);
        const bool has_counts = D->counts != nullptr, has_mask = has_counts && D->mask;
        const uint32_t nR = (uint32_t)D->nR;
        uint32_t total = 0;
        if (has_bkg) total += (uint32_t)kFoldTPX * 4u * nR;
        constexpr bool cmp = COMPACT;
        if (has_counts) total += (uint32_t)kFoldTPX * (cmp ? 2u : 4u) * nR;
        if (has_mask) total += (cmp ? 16u : (uint32_t)kFoldTPX) * nR;
        mbar_arrive_expect_tx(&bars[1], total);
        if (has_bkg) tma_load_2d(bs, tm + 128, p0, 0, &bars[1], pol_stream);
        if (has_counts) tma_load_2d(cs, tm + 256, p0, 0, &bars[1], pol_stream);
        if (has_mask) tma_load_2d(ms, tm + 384, cmp ? p0 >> 3 : p0, 0, &bars[1], pol_stream);  // 16 bytes of bits per row
      }
      return;
    }
    // copy index i = 3 r + which, spread over (lane, warp) like the exposure rows
    const int nR = D->nR;
    const float* bkg = D->background;
    const float* cnt = D->counts;
    const uint8_t* msk = D->mask;
    const size_t PP = (size_t)D->P;
    constexpr bool cmp = COMPACT;
    for (int i = (tid & 31) * (kFoldThreads / 32) + (tid >> 5); i < 3 * nR; i += kFoldThreads) {
      const int r = i / 3, which = i - 3 * r;
      const size_t off = (size_t)r * PP + p0;
      if (which == 0 && has_bkg) bulk_g2s_hint(bs + r * kFoldTPX, bkg + off, row, &bars[1], pol_stream);
      if (!cmp) {
        if (which == 1 && cnt) bulk_g2s_hint(cs + r * kFoldTPX, cnt + off, row, &bars[1], pol_stream);
        if (which == 2 && cnt && msk) bulk_g2s_hint(ms + r * kFoldTPX, msk + off, (uint32_t)npx, &bars[1], pol_stream);
      } else {
        if (which == 1 && cnt)
          bulk_g2s_hint(reinterpret_cast<unsigned short*>(cs) + r * kFoldTPX,
                        reinterpret_cast<const unsigned short*>(cnt) + off, row / 2u, &bars[1], pol_stream);
        if (which == 2 && cnt && msk)
          bulk_g2s_hint(ms + r * 16, msk + (size_t)r * D->mask_row_bytes + (p0 >> 3), 16u, &bars[1], pol_stream);
      }
    }
  };

  // One mbarrier per descriptor buffer: a buffer's next phase (the item after the next) is requested at the top of
  // the next item, after the end-of-item barrier behind which every thread has tested the current phase -- so the
  // request can go out at the very top of an item.  (With one barrier for both buffers a request issued before the
  // item's first CTA barrier could complete two phases ahead of a slow thread, whose parity wait then never returns.)
  auto fetch_desc = [&](int it, uint32_t buf) {  // one thread: one bulk copy of work item `it`'s descriptor
    if (tid != kClaimTid) return;
    mbar_arrive_expect_tx(&bars[2 + buf], (uint32_t)DL.bytes);
    bulk_g2s(desc_s + buf * DL.bytes, P.desc + (size_t)it * DL.bytes, (uint32_t)DL.bytes, &bars[2 + buf]);
  };

  // The per-step numbers of an item -- the flux vectors of its register-resident sources (zero rows beyond nl) and
  // its background factors, both K0 output in global memory (L2 / L1 resident) -- are copied into shared memory by
  // all threads one item ahead, during the previous item's epilogue; two CTA barriers lie between the copy and the
  // first read.  The descriptor in buffer `buf` must have landed.
  const int dyn_len = kFoldLMax * P.smem_nT + P.smem_nRp;
  auto load_dyn = [&](uint32_t buf, int i) -> double {  // element i of the item whose descriptor is in `buf`
    const unsigned char* dsc = desc_s + buf * DL.bytes;
    const int4 head = *reinterpret_cast<const int4*>(dsc);
    const InstDev* inst = reinterpret_cast<const InstDev*>(dsc + DL.inst);
    const int l = (i >= P.smem_nT) + (i >= 2 * P.smem_nT) + (i >= 3 * P.smem_nT) + (i >= 4 * P.smem_nT);
    const int t = i - l * P.smem_nT;
    if (l == kFoldLMax) return __ldg(P.bkgfac + (size_t)head.z * P.smem_nRp + t);
    if (l < head.x) return __ldg(P.flux + inst[l].flux_off + t);  // only t < the dataset's own nT is used
    return 0.0;
  };
  auto stage_dyn = [&](uint32_t buf) {
    for (int i = tid; i < dyn_len; i += kFoldThreads) dyn_s[buf * dyn_len + i] = load_dyn(buf, i);
  };

  const int n_work = P.work ? *P.n_work : n_items;
  auto item_at = [&](int pos) -> int { return pos < n_work ? (P.work ? __ldg(P.work + pos) : pos) : n_items; };
  auto desc_nl = [&](int pos) -> int {  // number of register-resident sources of the item at worklist position pos
    return __ldg(reinterpret_cast<const int*>(P.desc + (size_t)pos * DL.bytes));
  };
  const int* zero_i = reinterpret_cast<const int*>(P.zeros);
  // (Re)load the dataset table entry, its band-compressed edisp matrices and band table.  Block uniform.
  auto load_dataset = [&](int tile) {
    const int* src = reinterpret_cast<const int*>(P.ds + P.tile_ds[tile]);
    if (tid < (int)(sizeof(DatasetDev) / 4)) reinterpret_cast<int*>(&Ds)[tid] = src[tid];
    __syncthreads();
    const double* pdfc = Ds.pdfc;
    for (int i = tid; i < Ds.pdfc_rows * 8; i += kFoldThreads) pdf_s[i] = pdfc[i];
    const int nb = min(Ds.n_groups * (Ds.nRp >> 3) * 4, kBandSmem);
    const int* band = Ds.band;
    for (int i = tid; i < nb; i += kFoldThreads) band_s[i] = band[i];
  };

  // Per-thread view of an item's sources: this thread's pixel quad and, per overlapping source, a pointer into its
  // cutout; outside the cutout the pointer targets a zero word with stride 0, so phase 1 is branch free and its loads
  // can all be in flight.
  const int q = tid & 31, slot = tid >> 5;   // phase 1: 32 pixel quads x 8 energy slots
  const int pp = tid & 63, cslot = tid >> 6;  // phase 2: 64 pixel pairs x 4 reco-chunk slots
  // pointer / plane stride of source l of the descriptor in buffer `buf` at this thread's quad of item `which`
  auto quad_view = [&](uint32_t buf, int which, const float* (&cptr)[kFoldLMax], int (&pstr)[kFoldLMax],
                       int (&grp)[kFoldLMax], bool& qlive, int& qy, int& qx, int lmax) -> const DatasetDev* {
    int p0, npx;
    const DatasetDev* D = geom(which, p0, npx);
    const unsigned char* dsc = desc_s + buf * DL.bytes;
    const int nl = reinterpret_cast<const int*>(dsc)[0];
    const InstDev* slist = reinterpret_cast<const InstDev*>(dsc + DL.inst);
    const int nx = D->nx;
    const int pq = p0 + 4 * q;
    qlive = 4 * q < npx;
    qy = pq / nx;
    qx = pq - qy * nx;
#pragma unroll
    for (int l = 0; l < kFoldLMax; ++l) {
      cptr[l] = P.zeros;
      pstr[l] = 0;
      grp[l] = 0;
      if (l < lmax && l < nl && qlive) {
        const InstDev& in = slist[l];
        const int ly = qy - in.y0, lx = qx - in.x0a;
        grp[l] = in.group;
        if (ly >= 0 && ly < in.cy && lx >= 0 && lx < in.pitch && !(GPY_ABLATE & 1)) {
          cptr[l] = in.conv + ly * in.pitch + lx;
          pstr[l] = in.plane_stride;
        }
      }
    }
    return D;
  };
  // Items are claimed two ahead: `item` is being computed, `next` has its copies in flight, and the claim of the one
  // after (an atomic on a global counter: ~1 us round trip) is issued at the top of an item and consumed at its end.
  int w = blockIdx.x;
  int item = item_at(w);
  if (item >= n_items) return;
  int wn, next;
  if (!DYNAMIC) {
    wn = w + gridDim.x;
    next = item_at(wn);
  } else {
    if (tid == kClaimTid) {
      const int c = gridDim.x + atomicAdd(P.counter, 1u);
      nxt_s = make_int2(c, item_at(c));
    }
  }
  load_dataset(DYNAMIC ? item : item / P.B);
  __syncthreads();  // Ds, nxt_s
  if (DYNAMIC) {
    const int2 v = nxt_s;
    wn = v.x;
    next = v.y;
  }
  {
    const bool expo_first = desc_nl(w) > 0;  // source-free tiles do not fetch their exposure
    if (expo_first) arm_expo(item);
    arm_epi(item);
    fetch_desc(w, 0);
    __syncthreads();
    if (expo_first) issue_expo(item);
    issue_epi(item);
    mbar_wait(&bars[2], 0);
    stage_dyn(0);
    __syncthreads();
  }
  uint32_t it = 0;       // items done by this CTA: descriptor buffer it & 1, its mbarrier phase (it >> 1) & 1
  uint32_t eparity = 0;  // phase of the exposure barrier: flips on items that have sources only
#ifdef GPY_FOLD_TIMING
  long long tacc[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  long long tlast = clock64();
  long long cls_cyc[4] = {0, 0, 0, 0};
  int cls_n[4] = {0, 0, 0, 0};
#define GPY_TICK(i) do { if (tid == 0) { long long now_ = clock64(); tacc[i] += now_ - tlast; tlast = now_; } } while (0)
#else
#define GPY_TICK(i) do { } while (0)
#endif

  for (;; ++it) {
    const uint32_t parity = it & 1u;          // descriptor / per-step-number buffer of this item
    const uint32_t dphase = (it >> 1) & 1u;   // phase of that buffer's mbarrier
    const int b = DYNAMIC ? 0 : item % P.B;
    const int tile = DYNAMIC ? item : item / P.B;
    // claim the item after the next (consumed at the end of this item) and look up whether the next one has sources
    // (nothing may depend on the two results here: their consumers sit behind the phases, so the ~1 us round trips
    // are off the critical path)
    int wnn = 0, nn = n_items;
    unsigned claim_raw = 0;
    if (!DYNAMIC) {
      wnn = wn + gridDim.x;
      nn = item_at(wnn);
    } else if (tid == kClaimTid) {
      claim_raw = atomicAdd(P.counter, 1u);
    }
    const int nl_next = __ldg(next < n_items ? reinterpret_cast<const int*>(P.desc + (size_t)wn * DL.bytes) : zero_i);
#define expo_next (nl_next > 0)
    if (next < n_items) fetch_desc(wn, parity ^ 1u);
    if (tile < Ds.tile_begin || tile >= Ds.tile_begin + Ds.n_tiles) {  // block uniform: first item of another dataset
      __syncthreads();  // nobody still reads the old entry (issue_epi of this item at the end of the last one)
      load_dataset(tile);
      __syncthreads();
    }
    const int nT = Ds.nT, nR = Ds.nR, nRp = Ds.nRp, G = Ds.n_groups, nx = Ds.nx;
    const int p0 = (tile - Ds.tile_begin) * kFoldTPX;
    const int npx = min(kFoldTPX, (int)Ds.P - p0);
    const bool has_bkg = Ds.background && P.include_background, has_counts = Ds.counts != nullptr,
               has_mask = Ds.mask != nullptr;
    constexpr bool cmp = COMPACT;
    const int nchunks = nRp >> 3;
    const int* band = G * nchunks * 4 <= kBandSmem ? band_s : Ds.band;
    // work descriptor of this item (requested at the top of the previous item)
    mbar_wait(&bars[2 + parity], dphase);
    const unsigned char* dsc = desc_s + parity * DL.bytes;
    const int nl = reinterpret_cast<const int*>(dsc)[0], nover = reinterpret_cast<const int*>(dsc)[1];
    const InstDev* slist = reinterpret_cast<const InstDev*>(dsc + DL.inst);
    const int* over = reinterpret_cast<const int*>(dsc + DL.over);
    const double* flux_s = dyn_s + parity * dyn_len;              // [kFoldLMax][smem_nT]
    const double* bkgfac = flux_s + kFoldLMax * P.smem_nT;        // [smem_nRp]
    const int fstride = P.smem_nT;
    GPY_TICK(0);
    GPY_TICK(1);
#ifdef GPY_FOLD_TIMING
    const long long t_item0 = clock64();
#endif
    const float* cptr[kFoldLMax];
    int pstr[kFoldLMax], grp[kFoldLMax], qy, qx;
    bool qlive;
    quad_view(parity, item, cptr, pstr, grp, qlive, qy, qx, kFoldLMax);
    double stat = 0.0;
    for (int cbase = 0; cbase < nchunks; cbase += 4) {  // one pass for nR <= 32
      const int c = cbase + cslot;
      double acc0[8], acc1[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc0[j] = acc1[j] = 0.0;

      // No source reaches this tile: npred = background only; its exposure was never requested, so the buffer is
      // free for the next item's (armed here, issued behind the first barrier of the item).
      if (nl == 0 && cbase == 0 && expo_next) arm_expo(next);
      for (int t0 = 0; t0 < (nl > 0 ? nT : 0); t0 += kTmaTCH) {
        GPY_TICK(2);
        if (cbase == 0 && t0 == 0) mbar_wait(&bars[0], eparity);  // exposure tile has landed
        GPY_TICK(3);
        // ---- phase 1: v[g][t][p] = 1e4 * exposure[t,p] * sum_s F_s[t] * conv_s[t,p] ----
        for (int g = 0; g < G; ++g) {
          if (nl <= 2) {
            // Common case (CTA uniform): at most two sources reach the tile.  All eight loads of the pass are issued
            // before the first use, so the pass pays one memory latency instead of one per true-energy slot.
            float4 c0[kTmaTCH / 8], c1[kTmaTCH / 8];
#pragma unroll
            for (int tt = 0; tt < kTmaTCH / 8; ++tt) {
              const int tc = min(t0 + slot + 8 * tt, nT - 1);  // rows past the axis are loaded but not used
              c0[tt] = ldg_f4_hint(cptr[0] + (size_t)tc * pstr[0], pol_keep);
              c1[tt] = ldg_f4_hint(cptr[1] + (size_t)tc * pstr[1], pol_keep);
            }
#pragma unroll
            for (int tt = 0; tt < kTmaTCH / 8; ++tt) {
              const int tl = slot + 8 * tt, t = t0 + tl;
              double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
              if (t < nT && qlive) {
                const double F0 = grp[0] == g ? flux_s[t] : 0.0;
                const double F1 = grp[1] == g ? flux_s[fstride + t] : 0.0;
                s0 = fma(F1, (double)c1[tt].x, fma(F0, (double)c0[tt].x, s0));
                s1 = fma(F1, (double)c1[tt].y, fma(F0, (double)c0[tt].y, s1));
                s2 = fma(F1, (double)c1[tt].z, fma(F0, (double)c0[tt].z, s2));
                s3 = fma(F1, (double)c1[tt].w, fma(F0, (double)c0[tt].w, s3));
                const float4 e = *reinterpret_cast<const float4*>(es + t * kFoldTPX + 4 * q);
                s0 = (s0 * (double)e.x) * 1e4;
                s1 = (s1 * (double)e.y) * 1e4;
                s2 = (s2 * (double)e.z) * 1e4;
                s3 = (s3 * (double)e.w) * 1e4;
              }
              double* dst = vs + ((size_t)g * kTmaTCH + tl) * kFoldTPX + 4 * q;
              *reinterpret_cast<double2*>(dst) = make_double2(s0, s1);
              *reinterpret_cast<double2*>(dst + 2) = make_double2(s2, s3);
            }
            continue;
          }
          float4 cg[2][kFoldLMax];
#pragma unroll
          for (int tt = 0; tt < kTmaTCH / 8; ++tt) {
            const int tl = slot + 8 * tt, t = t0 + tl;
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
            // three or four sources: the loads of two slots (eight) are in flight together
            if ((tt & 1) == 0) {
#pragma unroll
              for (int u = 0; u < 2; ++u) {
                const int tc = min(t + 8 * u, nT - 1);
#pragma unroll
                for (int l = 0; l < kFoldLMax; ++l) cg[u][l] = ldg_f4_hint(cptr[l] + (size_t)tc * pstr[l], pol_keep);
              }
            }
            if (t < nT && qlive) {
#pragma unroll
              for (int l = 0; l < kFoldLMax; ++l) {
                const float4 cv = cg[tt & 1][l];
                const double F = grp[l] == g ? flux_s[l * fstride + t] : 0.0;
                s0 = fma(F, (double)cv.x, s0);
                s1 = fma(F, (double)cv.y, s1);
                s2 = fma(F, (double)cv.z, s2);
                s3 = fma(F, (double)cv.w, s3);
              }
              const float4 e = *reinterpret_cast<const float4*>(es + t * kFoldTPX + 4 * q);
              s0 = (s0 * (double)e.x) * 1e4;
              s1 = (s1 * (double)e.y) * 1e4;
              s2 = (s2 * (double)e.z) * 1e4;
              s3 = (s3 * (double)e.w) * 1e4;
            }
            double* dst = vs + ((size_t)g * kTmaTCH + tl) * kFoldTPX + 4 * q;
            *reinterpret_cast<double2*>(dst) = make_double2(s0, s1);
            *reinterpret_cast<double2*>(dst + 2) = make_double2(s2, s3);
          }
          // Further sources on this tile (dense fields: more than kFoldLMax cutouts overlap): one source at a time,
          // its loads of the pass' four slots in flight together, added to the values this thread just stored.
          if (qlive) {
            for (int o = 0; o < nover; ++o) {
              const InstDev& in = P.insts[over[o]];
              if (in.group != g) continue;
              const int ly = qy - in.y0, lx = qx - in.x0a;
              if (ly < 0 || ly >= in.cy || lx < 0 || lx >= in.pitch) continue;
              const float* base = in.conv + ly * in.pitch + lx;
              float4 co[kTmaTCH / 8];
#pragma unroll
              for (int tt = 0; tt < kTmaTCH / 8; ++tt)
                co[tt] = __ldg(reinterpret_cast<const float4*>(base + (size_t)min(t0 + slot + 8 * tt, nT - 1) * in.plane_stride));
#pragma unroll
              for (int tt = 0; tt < kTmaTCH / 8; ++tt) {
                const int tl = slot + 8 * tt, t = t0 + tl;
                if (t >= nT) continue;
                const double F = __ldg(P.flux + in.flux_off + t);
                const float4 e = *reinterpret_cast<const float4*>(es + t * kFoldTPX + 4 * q);
                double* dst = vs + ((size_t)g * kTmaTCH + tl) * kFoldTPX + 4 * q;
                double2 a = *reinterpret_cast<double2*>(dst), b2 = *reinterpret_cast<double2*>(dst + 2);
                a.x += ((F * (double)co[tt].x) * (double)e.x) * 1e4;
                a.y += ((F * (double)co[tt].y) * (double)e.y) * 1e4;
                b2.x += ((F * (double)co[tt].z) * (double)e.z) * 1e4;
                b2.y += ((F * (double)co[tt].w) * (double)e.w) * 1e4;
                *reinterpret_cast<double2*>(dst) = a;
                *reinterpret_cast<double2*>(dst + 2) = b2;
              }
            }
          }
        }
        GPY_TICK(4);
        const bool fetch_next = cbase + 4 >= nchunks && t0 + kTmaTCH >= nT && expo_next;
        if (fetch_next) arm_expo(next);
        __syncthreads();
        GPY_TICK(5);
        // exposure tile fully consumed: start fetching the next item's while phase 2 / the epilogue run
        if (fetch_next) issue_expo(next);
        GPY_TICK(6);
        // ---- phase 2: acc[r] += pdf[t][r] * v[t][p] over the non-zero band ----
        if (c < nchunks && !(GPY_ABLATE & 2)) {
          for (int g = 0; g < G; ++g) {
            const int4 bnd = *reinterpret_cast<const int4*>(band + (g * nchunks + c) * 4);
            const int blo = bnd.x, bhi = bnd.y;
            const int tlo = max(blo, t0), thi = min(bhi, t0 + kTmaTCH);
            const double* vrow = vs + (size_t)g * kTmaTCH * kFoldTPX + 2 * pp;
            const double* prow = pdf_s + (bnd.z - blo) * 8;  // band-compressed rows of this chunk
            for (int t = tlo; t < thi; ++t) {
              const double2 v = *reinterpret_cast<const double2*>(vrow + (t - t0) * kFoldTPX);
              const double2 m0 = *reinterpret_cast<const double2*>(prow + t * 8);
              const double2 m1 = *reinterpret_cast<const double2*>(prow + t * 8 + 2);
              const double2 m2 = *reinterpret_cast<const double2*>(prow + t * 8 + 4);
              const double2 m3 = *reinterpret_cast<const double2*>(prow + t * 8 + 6);
              const double m[8] = {m0.x, m0.y, m1.x, m1.y, m2.x, m2.y, m3.x, m3.y};
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                acc0[j] = fma(v.x, m[j], acc0[j]);
                acc1[j] = fma(v.y, m[j], acc1[j]);
              }
            }
          }
        }
        GPY_TICK(7);
        __syncthreads();
        GPY_TICK(5);
      }

      // ---- epilogue for reco chunk c: background, clip, Cash (inputs staged in shared memory) ----
      GPY_TICK(2);
      if (cbase == 0) mbar_wait(&bars[1], parity);
      GPY_TICK(8);
      // mu = acc + background model, clipped, parked in the (now free) v buffer as [r - r0][128]; the Cash terms
      // are then computed in a second pass that deals the bins round-robin over all 256 threads, so the warps
      // that own the low-energy reco chunks (where every bin has counts and needs a log) do not become the
      // critical path of the item.
      const int r0 = 8 * cbase;
      double* mus = vs;
      if (c < nchunks && 2 * pp < npx) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int r = 8 * c + j;
          if (r >= nR) break;
          double mu0 = acc0[j], mu1 = acc1[j];
          if (P.include_background) {
            if (has_bkg) {
              const float2 bg = *reinterpret_cast<const float2*>(bs + r * kFoldTPX + 2 * pp);
              mu0 += (double)bg.x * bkgfac[r];
              mu1 += (double)bg.y * bkgfac[r];
            }
            if (mu0 < 0.0) mu0 = 0.0;  // npred_total.data[npred_total.data < 0.0] = 0 (map.py:787)
            if (mu1 < 0.0) mu1 = 0.0;
          }
          *reinterpret_cast<double2*>(mus + (r - r0) * kFoldTPX + 2 * pp) = make_double2(mu0, mu1);
        }
      }
      __syncthreads();
      GPY_TICK(12);
      if (nl == 0 && cbase == 0 && expo_next) issue_expo(next);
      // next item's per-step numbers: the loads are issued before the Cash pass and stored to shared memory after
      // it (few registers are live here), so the pass covers their latency; the end-of-item barrier publishes them
      const bool stage_next = cbase + 4 >= nchunks && next < n_items;
      double dyn0 = 0.0, dyn1 = 0.0;
      if (stage_next) {
        mbar_wait(&bars[2 + (parity ^ 1u)], ((it + 1u) >> 1) & 1u);  // the next descriptor, requested at the top of this item
        if (tid < dyn_len) dyn0 = load_dyn(parity ^ 1u, tid);
        if (tid + kFoldThreads < dyn_len) dyn1 = load_dyn(parity ^ 1u, tid + kFoldThreads);
      }
      GPY_TICK(13);
      {
        // Two neighbouring pixels per thread and iteration (npx is even: P % 16 == 0 on this path): one 128-bit load
        // of mu, one 64-bit load of the counts, one 16-bit load of the mask.  Most bins of a gamma-ray cube hold no
        // count, so the log is the rare branch and the pass is bound by its loads and index arithmetic.
        const int nr = min(nR - r0, 32);
        const double log_trunc = P.log_trunc;
        // The trip count is warp uniform (nr * 64 items, 32 consecutive ones per warp), so the warp can vote: the two
        // logs run only when some lane holds a count, as straight-line code for both pixels at once (fast_log: no
        // data-dependent branch, the two dependency chains interleave).  Separate accumulators for the mu and the
        // n log mu terms halve the fp64 add chain.
        double stat_mu = 0.0, stat_nl = 0.0;
        for (int i = tid; i < ((GPY_ABLATE & 4) ? 0 : nr * (kFoldTPX / 2)); i += kFoldThreads) {
          const int rr = i >> 6, lp = (i & (kFoldTPX / 2 - 1)) * 2;  // kFoldTPX == 128
          const bool live = lp < npx;
          const int r = r0 + rr;
          const double2 mu2 = *reinterpret_cast<const double2*>(mus + rr * kFoldTPX + lp);
          if (WRITE_NPRED && live)
            *reinterpret_cast<double2*>(P.npred_out + (size_t)r * (size_t)Ds.P + p0 + lp) = mu2;
          if (has_counts) {
            float2 n2;
            unsigned m2;
            if (!cmp) {
              n2 = *reinterpret_cast<const float2*>(cs + r * kFoldTPX + lp);
              m2 = has_mask ? *reinterpret_cast<const unsigned short*>(ms + r * kFoldTPX + lp) : 0x0101u;
            } else {  // uint16 counts, one mask bit per bin (lp is even: both bits sit in one byte)
              const unsigned nn = *reinterpret_cast<const unsigned*>(reinterpret_cast<const unsigned short*>(cs) + r * kFoldTPX + lp);
              n2 = make_float2((float)(nn & 0xffffu), (float)(nn >> 16));
              const unsigned mb = has_mask ? (unsigned)ms[r * 16 + (lp >> 3)] >> (lp & 7) : 3u;
              m2 = (mb & 1u) | ((mb & 2u) << 7);
            }
            const bool on0 = live && (m2 & 0xffu), on1 = live && (m2 >> 8);
            const bool big0 = mu2.x > kTrunc, big1 = mu2.y > kTrunc;
            const double a0 = big0 ? mu2.x : kTrunc, a1 = big1 ? mu2.y : kTrunc;
            stat_mu += (on0 ? a0 : 0.0) + (on1 ? a1 : 0.0);
            const bool need0 = on0 && n2.x > 0.f, need1 = on1 && n2.y > 0.f;
            if (__any_sync(0xffffffffu, need0 || need1)) {
              double l0 = fast_log(a0, logtab_s), l1 = fast_log(a1, logtab_s);
              l0 = big0 ? (a0 <= 1.7976931348623157e308 ? l0 : a0) : log_trunc;  // log(inf) = inf
              l1 = big1 ? (a1 <= 1.7976931348623157e308 ? l1 : a1) : log_trunc;
              stat_nl += (need0 ? (double)n2.x * l0 : 0.0) + (need1 ? (double)n2.y * l1 : 0.0);
            }
          }
        }
        stat += stat_mu - stat_nl;
      }
      GPY_TICK(14);
      if (stage_next) {
        double* dst = dyn_s + (parity ^ 1u) * dyn_len;
        if (tid < dyn_len) dst[tid] = dyn0;
        if (tid + kFoldThreads < dyn_len) dst[tid + kFoldThreads] = dyn1;
        for (int i = tid + 2 * kFoldThreads; i < dyn_len; i += kFoldThreads) dst[i] = load_dyn(parity ^ 1u, i);  // long axes
      }
      if (cbase + 4 < nchunks) __syncthreads();  // the next sweep's phase 1 rewrites the v buffer
    }
    GPY_TICK(9);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) stat += __shfl_xor_sync(0xffffffffu, stat, o);
    if ((tid & 31) == 0) wsum[tid >> 5] = stat;
    if (next < n_items) arm_epi(next);
    if (DYNAMIC && tid == kClaimTid) {
      const int c = gridDim.x + claim_raw;
      nxt_s = make_int2(c, item_at(c));
    }
    __syncthreads();  // every warp is past its epilogue reads; list / bkgfac reuse in the next item
    if (tid == 0)     // fixed summation order; no warp can rewrite wsum before passing a barrier with thread 0
      P.partials[(size_t)b * P.n_tiles + tile] =
          ((wsum[0] + wsum[1]) + (wsum[2] + wsum[3])) + ((wsum[4] + wsum[5]) + (wsum[6] + wsum[7]));
    GPY_TICK(10);
    if (next < n_items) issue_epi(next);
    if (nl > 0) eparity ^= 1u;
    if (DYNAMIC) {  // no thread can pass the next end-of-item barrier (where thread 0 rewrites nxt_s) before this read
      const int2 v = nxt_s;
      wnn = v.x;
      nn = v.y;
    }
    const bool done = next >= n_items;
    item = next;
    w = wn;
    next = nn;
    wn = wnn;
    GPY_TICK(11);
#ifdef GPY_FOLD_TIMING
    {
      const int k = nl == 0 ? 0 : nl <= 2 ? 1 : nover == 0 ? 2 : 3;
      cls_cyc[k] += clock64() - t_item0;
      cls_n[k]++;
    }
#endif
    if (done) break;
  }
#undef expo_next
#ifdef GPY_FOLD_TIMING
  if (tid == 0 && (blockIdx.x == 0 || blockIdx.x == 100))
    printf("[fold timing] block %d: list %lld | sync-after-list %lld | misc %lld | expo wait %lld | phase1 %lld | phase syncs %lld | issue expo %lld | phase2 %lld | epi wait %lld | park+barrier %lld | stage-next issue %lld | cash %lld | dyn store %lld | end barrier %lld | issue epi %lld\n",
           blockIdx.x, tacc[0], tacc[1], tacc[2], tacc[3], tacc[4], tacc[5], tacc[6], tacc[7], tacc[8], tacc[12], tacc[13], tacc[14], tacc[9], tacc[10], tacc[11]);
  if (tid == 0 && (blockIdx.x == 0 || blockIdx.x == 100 || blockIdx.x == 201))
    printf("[fold classes] block %d: no-source %d items %lld cyc | 1-2 sources %d items %lld cyc | 3-4 sources %d items %lld cyc | >4 sources %d items %lld cyc\n",
           blockIdx.x, cls_n[0], cls_cyc[0], cls_n[1], cls_cyc[1], cls_n[2], cls_cyc[2], cls_n[3], cls_cyc[3]);
#endif
}

// Second stage: stat[b, d] = 2 * sum of the tile partials of dataset d, fixed summation order (thread t adds tiles
// t, t + 512, ..., then a fixed tree): 512 threads keep the dependent chain of a C3-sized dataset at 16 loads.
constexpr int kReduceThreads = 512;
__global__ void __launch_bounds__(kReduceThreads)
reduce_stat_kernel(const double* __restrict__ partials, const DatasetDev* __restrict__ ds, int n_tiles,
                   int n_datasets, int B, const unsigned char* __restrict__ dup, double* __restrict__ stat) {
  __shared__ double sm[kReduceThreads];
  const int b = blockIdx.x, d = blockIdx.y;
  const int begin = ds[d].tile_begin, n = ds[d].n_tiles;
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += kReduceThreads) {
    const int tile = begin + i;
    // items identical to vector 0 on this tile were not evaluated: read the partial of (tile, 0)
    const int bb = (dup && dup[(size_t)tile * B + b]) ? 0 : b;
    s += partials[(size_t)bb * n_tiles + tile];
  }
  sm[threadIdx.x] = s;
  __syncthreads();
  for (int o = kReduceThreads / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) stat[(size_t)b * n_datasets + d] = 2.0 * sm[0];
}

// ----------------------------------------------------------------------------------------------
// Joint fits over ranks (datasets/actors.py:84-87, 168-173: the sum of the per-actor statistics): one-shot all-reduce
// of the B local statistics over NVLink peer memory.  Every rank owns a buffer that all ranks have mapped
// (symmetric memory); layout per parity p in {0, 1} and source rank r:  double vals[cap], then one 64-bit flag.
// A launch (one per evaluation, the same count on every rank) takes the next sequence number from a device-local
// counter, stores its own values into slot (seq & 1, rank) of EVERY rank's buffer, publishes them with a
// system-scope release store of seq into the slot's flag, then waits on its own buffer until all flags of this
// parity show seq and adds the values in rank order (the same sum on every rank, bit for bit).  Two parities
// suffice: nobody can write seq + 2 before everybody has read seq (they all wait for everybody's seq + 1).
// ----------------------------------------------------------------------------------------------
constexpr int kPeerMaxWorld = 16;
struct PeerReduceParams {
  double* bufs[kPeerMaxWorld];   // this process' mappings of the ranks' buffers
  unsigned long long* seq;       // device-local sequence counter
  int rank, world, cap;          // cap: values per slot
  int n_datasets;
};

__device__ __forceinline__ size_t peer_slot(const PeerReduceParams& Q, int parity, int r) {
  return ((size_t)parity * Q.world + r) * (size_t)(Q.cap + 1);
}

__global__ void __launch_bounds__(256)
peer_allreduce_kernel(const PeerReduceParams Q, const double* __restrict__ stat /* [B, n_datasets] */, int B,
                      double* __restrict__ out /* [B] */) {
  __shared__ unsigned long long seq_s;
  if (threadIdx.x == 0) seq_s = *Q.seq + 1ull;
  __syncthreads();
  const unsigned long long seq = seq_s;
  const int parity = (int)(seq & 1ull);
  // local totals in dataset order (Datasets.stat_sum), stored into every rank's slot of this rank
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    double t = 0.0;
    for (int d = 0; d < Q.n_datasets; ++d) t += stat[(size_t)b * Q.n_datasets + d];
    for (int q = 0; q < Q.world; ++q) Q.bufs[q][peer_slot(Q, parity, Q.rank) + b] = t;
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < Q.world) {
    unsigned long long* flag = reinterpret_cast<unsigned long long*>(Q.bufs[threadIdx.x] + peer_slot(Q, parity, Q.rank) + Q.cap);
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag), "l"(seq) : "memory");
  }
  // wait for every rank's values of this sequence number (bounded: a rank that never launches must not hang the GPU)
  __shared__ int bad;
  if (threadIdx.x == 0) bad = 0;
  __syncthreads();
  if (threadIdx.x < Q.world) {
    const unsigned long long* flag =
        reinterpret_cast<const unsigned long long*>(Q.bufs[Q.rank] + peer_slot(Q, parity, threadIdx.x) + Q.cap);
    unsigned long long t0, now, seen;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(flag) : "memory");
      if (seen >= seq) break;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (now - t0 > 5000000000ull) {  // 5 s
        bad = 1;
        break;
      }
      __nanosleep(20);
    }
  }
  __syncthreads();
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    double t = 0.0;
    for (int r = 0; r < Q.world; ++r) t += Q.bufs[Q.rank][peer_slot(Q, parity, r) + b];
    out[b] = bad ? __longlong_as_double(0x7ff8000000000000LL) : t;
  }
  if (threadIdx.x == 0) *Q.seq = seq;
}

// ----------------------------------------------------------------------------------------------
// Stand-alone Cash sums (stats/fit_statistics_cython.pyx:17-85) over already-masked 1-D fp64 arrays.
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
cash_sum_kernel(const double* __restrict__ counts, const double* __restrict__ npred,
                const double* __restrict__ weight, long long n, double log_trunc,
                double* __restrict__ partials) {
  __shared__ double sm[256];
  double s = 0.0;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    double npr = npred[i], c = counts[i], lognpr = 0.0;
    if (npr > kTrunc) {
      if (c > 0.0) lognpr = log(npr);
    } else {
      npr = kTrunc;
      lognpr = log_trunc;
    }
    if (weight) {
      double w = weight[i];
      if (w > 0.0) {
        s += w * npr;
        if (c > 0.0) s -= w * c * lognpr;
      }
    } else {
      s += npr;
      if (c > 0.0) s -= c * lognpr;
    }
  }
  sm[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partials[blockIdx.x] = sm[0];
}

// ----------------------------------------------------------------------------------------------
// WStat (stats/fit_statistics.py:142-252, WStatFitStatistic :332-361): ON counts with a Poisson OFF measurement.
//   mu_bkg = (C + D) / (2 alpha (alpha + 1)),  C = alpha (n_on + n_off) - (1 + alpha) mu_sig,
//                                               D = sqrt(C^2 + 4 alpha (alpha + 1) n_off mu_sig)      (:221-236)
//   stat = 2 (mu_sig + (1 + alpha) mu_bkg - [n_on != 0] n_on log(mu_sig + alpha mu_bkg) - [n_off != 0] n_off log(mu_bkg))
//          + 2 ([n_on != 0] -n_on (1 - log n_on) + [n_off != 0] -n_off (1 - log n_off))                (:142-218, 239-252)
// then np.nan_to_num per bin (:349) and the sum over the bins where `mask` is true.  Inputs are device arrays: ON / OFF
// counts and alpha float32 or float64 (TI), mu_sig float64 (the npred_signal cube), mask uint8 or null.
// Products the reference rounds separately are kept apart (__dmul_rn / __dadd_rn): no FMA contraction.
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ double wstat_bin(double n_on, double n_off, double alpha, double mu_sig) {
  const double a1 = __dadd_rn(alpha, 1.0);
  const double C = __dadd_rn(__dmul_rn(alpha, __dadd_rn(n_on, n_off)), -__dmul_rn(__dadd_rn(1.0, alpha), mu_sig));
  const double four = __dmul_rn(__dmul_rn(__dmul_rn(__dmul_rn(4.0, alpha), a1), n_off), mu_sig);
  const double D = sqrt(__dadd_rn(__dmul_rn(C, C), four));
  const double mu_bkg = __ddiv_rn(__dadd_rn(C, D), __dmul_rn(__dmul_rn(2.0, alpha), a1));
  const double term1 = __dadd_rn(mu_sig, __dmul_rn(__dadd_rn(1.0, alpha), mu_bkg));
  const double term2 = n_on == 0.0 ? 0.0 : __dmul_rn(-n_on, log(__dadd_rn(mu_sig, __dmul_rn(alpha, mu_bkg))));
  const double term3 = n_off == 0.0 ? 0.0 : __dmul_rn(-n_off, log(mu_bkg));
  double stat = __dmul_rn(2.0, __dadd_rn(__dadd_rn(term1, term2), term3));
  double gof = 0.0;
  gof = __dadd_rn(gof, n_on == 0.0 ? 0.0 : __dmul_rn(-n_on, __dadd_rn(1.0, -log(n_on))));
  gof = __dadd_rn(gof, n_off == 0.0 ? 0.0 : __dmul_rn(-n_off, __dadd_rn(1.0, -log(n_off))));
  stat = __dadd_rn(stat, __dmul_rn(2.0, gof));
  if (isnan(stat)) return 0.0;  // np.nan_to_num
  if (isinf(stat)) return stat > 0 ? 1.7976931348623157e308 : -1.7976931348623157e308;
  return stat;
}

template <typename TI>
__global__ void __launch_bounds__(256)
wstat_sum_kernel(const TI* __restrict__ n_on, const TI* __restrict__ n_off, const TI* __restrict__ alpha,
                 const double* __restrict__ mu_sig, const unsigned char* __restrict__ mask, long long n,
                 double* __restrict__ per_bin, double* __restrict__ partials) {
  __shared__ double sm[256];
  double s = 0.0;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const double v = wstat_bin((double)n_on[i], (double)n_off[i], (double)alpha[i], mu_sig[i]);
    if (per_bin) per_bin[i] = v;
    if (!mask || mask[i]) s += v;
  }
  sm[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partials[blockIdx.x] = sm[0];
}

__global__ void __launch_bounds__(256)
final_sum_kernel(const double* __restrict__ partials, int n, double scale, double* __restrict__ out) {
  __shared__ double sm[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) s += partials[i];
  sm[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = scale * sm[0];
}

}  // namespace gpy
