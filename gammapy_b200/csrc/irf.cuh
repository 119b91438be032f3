// IRF kernel extraction at a source position: the image stage of PSFMap.get_psf_kernel
// (irf/psf/map.py:296-337) on the device.  The host (extract_psf in gpy_b200.cu) interpolates the radial profile of
// every true-energy bin at the source position, finds the containment radii and the per-plane cut / oversampling
// factor; this kernel evaluates the profile on the oversampled kernel grid, block-sums, normalises every plane and
// writes it flipped, cropped to its own cut and padded -- the layout psf_conv_kernel consumes (ConvPlane).
#pragma once
#include "kernels.cuh"

namespace gpy {

struct PsfPlaneJob {  // one true-energy plane
  int n;       // odd side of this plane's cut image (pixels): round_up_to_odd(2 r_t / binsz)
  int f;       // oversampling factor of this plane (<= 4)
  int koff;    // offset of the plane in the flipped kernel buffer (ConvPlane::koff)
  int kw;      // padded row width of the plane there (ConvPlane::kw)
};

struct PsfExtractParams {
  const PsfPlaneJob* jobs;   // [n_planes]
  const double* profile;     // [n_planes][n_rad] psf value (sr-1) at the rad bin centres, at the source position
  const double* rad_center;  // [n_rad] deg
  int n_rad, n_planes;
  double binsz;              // deg, pixel size of the kernel geometry
  int proj;                  // gpy_projection of the dataset geometry: the kernel geometry is WcsGeom.create(skydir =
                             // its centre, proj) (geom.to_odd_npix, maps/wcs/geom.py:1106-1138), so a pixel's offset
                             // from the centre pixel IS its native coordinate and its separation from the centre does
                             // not depend on where on the sky the centre lies
  int pad_;
  double* raw;               // scratch [sum n^2] doubles: block sums before normalisation, plane t at raw_off[t]
  const int* raw_off;        // [n_planes]
  float* kflip;              // output planes (flipped, padded rows), zero initialised by the caller
  float* kernel;             // optional [n_planes][K][K] unflipped normalised kernel (inspection / tests) or null
  int K;
};

// One block per plane.  Pass 1: every thread block-sums the f x f sub-pixels of its pixels (profile linear in rad
// between the bin centres, extrapolated with the edge segments, clipped at 0: Map.interp_by_coord(method="linear",
// fill_value=None) + np.clip; rad = angular separation to the geometry's centre).  Pass 2: plane sum in a fixed order
// (PSFKernel.normalize, irf/psf/kernel.py:70-80) and the scaled values into the flipped layout.
__global__ void __launch_bounds__(256) psf_extract_kernel(const PsfExtractParams Q) {
  extern __shared__ double psf_sm[];  // [n_rad] profile, [n_rad] centres, [256] reduction
  const int t = blockIdx.x;
  const PsfPlaneJob J = Q.jobs[t];
  double* prof = psf_sm;
  double* cen = psf_sm + Q.n_rad;
  double* red = cen + Q.n_rad;
  for (int i = threadIdx.x; i < Q.n_rad; i += blockDim.x) {
    prof[i] = Q.profile[(size_t)t * Q.n_rad + i];
    cen[i] = Q.rad_center[i];
  }
  __syncthreads();
  const int n = J.n, f = J.f, m = n * f;
  const double step = Q.binsz / f;
  const double half = (m - 1) / 2.0;
  double* raw = Q.raw + Q.raw_off[t];
  double local = 0.0;
  for (int p = threadIdx.x; p < n * n; p += blockDim.x) {
    const int iy = p / n, ix = p - iy * n;
    double s = 0.0;
    for (int sy = 0; sy < f; ++sy) {
      const double y = __dmul_rn((double)(iy * f + sy) - half, step);
      for (int sx = 0; sx < f; ++sx) {
        const double x = -__dmul_rn((double)(ix * f + sx) - half, step);
        // CAR: native (phi, theta) = (x, y), separation from the native origin; TAN: the plane radius is cot(theta)
        const double rad = Q.proj == 1 ? atan(hypot(x, y) * kDeg) / kDeg
                                       : angular_separation(x * kDeg, y * kDeg, 0.0, 0.0) / kDeg;
        // np.searchsorted(centers, rad) - 1 clipped to [0, n_rad - 2]
        int lo = 0, hi = Q.n_rad;
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (cen[mid] < rad) lo = mid + 1; else hi = mid;
        }
        int idx = min(max(lo - 1, 0), Q.n_rad - 2);
        const double w = (rad - cen[idx]) / (cen[idx + 1] - cen[idx]);
        const double v = prof[idx] * (1.0 - w) + prof[idx + 1] * w;
        s += v > 0.0 ? v : 0.0;
      }
    }
    raw[p] = s;
    local += s;
  }
  red[threadIdx.x] = local;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  const double total = red[0];
  const int h = (n - 1) / 2, hK = (Q.K - 1) / 2;
  for (int p = threadIdx.x; p < n * n; p += blockDim.x) {
    const int iy = p / n, ix = p - iy * n;
    double v = raw[p] / total;
    if (!isfinite(v)) v = 0.0;  // np.nan_to_num(data / sum): an all-zero plane stays zero
    const float vf = (float)v;
    // flipped: kflip[i][j] = plane[2h - i][2h - j]
    Q.kflip[J.koff + (size_t)(2 * h - iy) * J.kw + (2 * h - ix)] = vf;
    if (Q.kernel) Q.kernel[((size_t)t * Q.K + (hK - h + iy)) * Q.K + (hK - h + ix)] = vf;
  }
}

}  // namespace gpy
