// gpy_b200.cu -- host runtime + C ABI of libgpy_b200.so (see include/gpy_b200.h).
//
// The runtime plays the role of gammapy's MapEvaluator bookkeeping (datasets/evaluator.py:128-243,
// 282-286, 465-481): per source it decides when the exposure cutout has to be re-centred, when the
// spatial model has to be re-integrated and re-convolved with the PSF (cached otherwise), and packs
// one launch of the fused fold + Cash kernel for all datasets and all parameter vectors of a call.
// All integer geometry (Cutout2D slices, oversampling factors) is computed here in IEEE doubles with
// the reference's operation order, because a different rounding shifts a cutout by one pixel.
//
// Reference citations are relative to /root/reference/gammapy.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>
#include <string>
#include <vector>

#include <cuda.h>

#include "../../include/gpy_b200.h"
#include "kernels.cuh"
#include "tsmap.cuh"
#include "irf.cuh"
#include "fold_t64.cuh"

namespace {

thread_local std::string g_last_error;

int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

#define CUDA_TRY(expr)                                                                     \
  do {                                                                                     \
    cudaError_t err__ = (expr);                                                            \
    if (err__ != cudaSuccess)                                                              \
      return fail(GPY_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(err__), \
                  __FILE__, __LINE__);                                                     \
  } while (0)

#define GPY_TRY(expr)          \
  do {                         \
    int rc__ = (expr);         \
    if (rc__ != GPY_OK) return rc__; \
  } while (0)

constexpr double kDegH = 0.017453292519943295;
constexpr double kCutoutMargin = 0.1;  // deg, datasets/evaluator.py:17
constexpr int kMaxOversampling = 200;  // modeling/models/spatial.py:54
constexpr int kSlotsPerSource = 6;     // cached (spatial params, cutout) -> PSF-convolved cube, kept between calls
constexpr int kSlotsHardCap = 96;      // beyond this the per-call extras are trimmed at the next call

// ---------------------------------------------------------------------------------------------
// host geometry: the integer rules of astropy Cutout2D / overlap_slices and gammapy WcsGeom.cutout
// ---------------------------------------------------------------------------------------------
struct GeomH {
  int nx, ny;
  double binsz, crval_lon, crpix_x, crpix_y;
  double crval_lat = 0.0;
  int proj = 0;  // gpy_projection: 0 CAR, 1 TAN, both about (crval_lon, crval_lat)
  int frame = 0;  // gpy_frame of the world coordinates
};
GeomH geom_from_desc(const gpy_geom_desc& g) {
  GeomH h{g.nx, g.ny, g.binsz, g.crval_lon, g.crpix_x, g.crpix_y};
  h.crval_lat = g.crval_lat;
  h.proj = g.proj;
  h.frame = g.frame;
  return h;
}

// ---- celestial frames: ICRS <-> Galactic as astropy defines it (see gpy_frame in the header) -------------------
struct FrameRotation {
  double icrs_to_gal[9];
  FrameRotation() {
    auto rot = [](double deg, char axis, double m[9]) {  // astropy rotation_matrix: passive rotation about an axis
      const double a = deg * M_PI / 180.0, c = std::cos(a), s = std::sin(a);
      const double x[9] = {1, 0, 0, 0, c, s, 0, -s, c}, y[9] = {c, 0, -s, 0, 1, 0, s, 0, c}, z[9] = {c, s, 0, -s, c, 0, 0, 0, 1};
      std::memcpy(m, axis == 'x' ? x : axis == 'y' ? y : z, sizeof(x));
    };
    auto mul = [](const double a[9], const double b[9], double out[9]) {
      double t[9];
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) t[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
      std::memcpy(out, t, sizeof(t));
    };
    double a[9], b[9], bias[9], gal[9];
    // ICRS -> FK5 (J2000): frame bias of USNO circular 179, eta0 = -19.9 mas, xi0 = 9.1 mas, da0 = -22.9 mas
    rot(19.9 / 3600000.0, 'x', a);
    rot(9.1 / 3600000.0, 'y', b);
    mul(a, b, bias);
    rot(-22.9 / 3600000.0, 'z', b);
    mul(bias, b, bias);
    // FK5 (J2000) -> Galactic: north galactic pole (192.8594812065348, 27.12825118085622) deg, lon0 122.9319185680026 deg
    rot(180.0 - 122.9319185680026, 'z', a);
    rot(90.0 - 27.12825118085622, 'y', b);
    mul(a, b, gal);
    rot(192.8594812065348, 'z', b);
    mul(gal, b, gal);
    mul(gal, bias, icrs_to_gal);
  }
};
const FrameRotation kFrames;

bool valid_frame(int f) { return f == GPY_FRAME_GALACTIC || f == GPY_FRAME_ICRS; }

// Matrix taking unit vectors of frame `from` to frame `to` (row major); false when the frames are the same.
bool frame_matrix(int from, int to, double m[9]) {
  if (from == to) return false;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      m[3 * i + j] = from == GPY_FRAME_ICRS ? kFrames.icrs_to_gal[3 * i + j] : kFrames.icrs_to_gal[3 * j + i];
  return true;
}

// (lon, lat) in deg of frame `from` expressed in frame `to`.
void frame_convert(int from, int to, double lon, double lat, double& lon_out, double& lat_out) {
  double m[9];
  if (!frame_matrix(from, to, m)) {
    lon_out = lon;
    lat_out = lat;
    return;
  }
  const double lo = lon * M_PI / 180.0, la = lat * M_PI / 180.0;
  const double x = std::cos(la) * std::cos(lo), y = std::cos(la) * std::sin(lo), z = std::sin(la);
  const double x2 = m[0] * x + m[1] * y + m[2] * z, y2 = m[3] * x + m[4] * y + m[5] * z, z2 = m[6] * x + m[7] * y + m[8] * z;
  lon_out = std::atan2(y2, x2) * 180.0 / M_PI;
  if (lon_out < 0) lon_out += 360.0;
  lat_out = std::atan2(z2, std::hypot(x2, y2)) * 180.0 / M_PI;
}

struct Rect {
  int x0 = 0, y0 = 0, cx = 0, cy = 0;
  bool operator==(const Rect& o) const { return x0 == o.x0 && y0 == o.y0 && cx == o.cx && cy == o.cy; }
};

enum CutStatus { CUT_OK = 0, CUT_NO_OVERLAP = 1, CUT_VALUE_ERROR = 2 };

double py_mod(double a, double b) {  // Python float %
  double m = std::fmod(a, b);
  if (m != 0.0 && ((b < 0) != (m < 0))) m += b;
  return m;
}

void world_to_pix(const GeomH& g, double lon, double lat, double& px, double& py) {
  // wcs_world2pix(lon, lat, 0): pix = img / cdelt + crpix - 1 (wcslib linx2p simple path)
  if (g.proj == 1) {  // TAN (gnomonic, FITS WCS paper II eq. 54 with the native pole at the reference point)
    const double dl = (lon - g.crval_lon) * kDegH, la = lat * kDegH, lp = g.crval_lat * kDegH;
    const double s = std::sin(la) * std::sin(lp) + std::cos(la) * std::cos(lp) * std::cos(dl);
    if (!(s > 0)) {  // behind the tangent plane
      px = py = std::numeric_limits<double>::quiet_NaN();
      return;
    }
    const double x = (std::cos(la) * std::sin(dl) / s) / kDegH;
    const double y = ((std::sin(la) * std::cos(lp) - std::cos(la) * std::sin(lp) * std::cos(dl)) / s) / kDegH;
    px = x / (-g.binsz) + g.crpix_x - 1.0;
    py = y / g.binsz + g.crpix_y - 1.0;
    return;
  }
  if (g.crval_lat != 0.0) {  // oblique plate carree (see WcsDev): tilt back by crval_lat, native (phi, theta) = (x, y)
    const double dl = (lon - g.crval_lon) * kDegH, la = lat * kDegH, l0 = g.crval_lat * kDegH;
    const double xs = std::cos(la) * std::cos(dl), ys = std::cos(la) * std::sin(dl), zs = std::sin(la);
    const double x = xs * std::cos(l0) + zs * std::sin(l0), z = -xs * std::sin(l0) + zs * std::cos(l0);
    px = (std::atan2(ys, x) / kDegH) / (-g.binsz) + g.crpix_x - 1.0;
    py = (std::atan2(z, std::hypot(x, ys)) / kDegH) / g.binsz + g.crpix_y - 1.0;
    return;
  }
  double dlon = py_mod(lon - g.crval_lon + 180.0, 360.0) - 180.0;
  px = dlon / (-g.binsz) + g.crpix_x - 1.0;
  py = lat / g.binsz + g.crpix_y - 1.0;
}

double round_up_to_odd(double f) {  // utils/array.py:110
  return std::floor(std::ceil(f) / 2.0) * 2.0 + 1.0;
}

// WcsGeom.cutout (maps/wcs/geom.py:886-930) -> parent slices, mode "trim"
CutStatus cutout_rect(const GeomH& g, double lon, double lat, double width_deg, bool odd_npix, Rect& out) {
  double width_npix = width_deg / g.binsz;
  // np.clip(x, min_npix=1, None); NaN propagates
  if (width_npix < 1.0) width_npix = 1.0;
  if (odd_npix) width_npix = round_up_to_odd(width_npix);
  double rounded = std::nearbyint(width_npix);  // np.round: half to even
  if (!std::isfinite(rounded)) return CUT_VALUE_ERROR;  // int(nan) raises ValueError
  long shape = (long)rounded;
  double px, py;
  world_to_pix(g, lon, lat, px, py);
  if (!std::isfinite(px) || !std::isfinite(py)) return CUT_VALUE_ERROR;
  const double pos[2] = {py, px};
  const int large[2] = {g.ny, g.nx};
  long emin[2], emax[2];
  for (int a = 0; a < 2; ++a) {
    emin[a] = (long)std::ceil(pos[a] - (double)shape / 2.0);
    emax[a] = (long)std::ceil(pos[a] + (double)shape / 2.0);
  }
  for (int a = 0; a < 2; ++a)
    if (emax[a] <= 0) return CUT_NO_OVERLAP;
  for (int a = 0; a < 2; ++a)
    if (emin[a] >= large[a]) return CUT_NO_OVERLAP;
  long y0 = std::max(0L, emin[0]), y1 = std::min((long)large[0], emax[0]);
  long x0 = std::max(0L, emin[1]), x1 = std::min((long)large[1], emax[1]);
  out.x0 = (int)x0;
  out.y0 = (int)y0;
  out.cx = (int)(x1 - x0);
  out.cy = (int)(y1 - y0);
  if (out.cx <= 0 || out.cy <= 0) return CUT_NO_OVERLAP;
  return CUT_OK;
}

GeomH sub_geom(const GeomH& g, const Rect& r) {  // Cutout2D.wcs: crpix -= origin
  GeomH s = g;
  s.nx = r.cx;
  s.ny = r.cy;
  s.crpix_x = g.crpix_x - r.x0;
  s.crpix_y = g.crpix_y - r.y0;
  return s;
}

double angular_separation_h(double lon1, double lat1, double lon2, double lat2) {
  double sdlon = std::sin(lon2 - lon1), cdlon = std::cos(lon2 - lon1);
  double slat1 = std::sin(lat1), slat2 = std::sin(lat2);
  double clat1 = std::cos(lat1), clat2 = std::cos(lat2);
  double num1 = clat2 * sdlon;
  double num2 = clat1 * slat2 - slat1 * clat2 * cdlon;
  double den = slat1 * slat2 + clat1 * clat2 * cdlon;
  return std::atan2(std::hypot(num1, num2), den);
}

int n_spatial_params(int type) { return type == GPY_SPATIAL_POINT ? 2 : type == GPY_SPATIAL_GAUSS ? 5 : 6; }
int n_spectral_params(int type) { return type == GPY_SPECTRAL_PL ? 3 : type == GPY_SPECTRAL_LP ? 4 : 5; }

double evaluation_radius(int type, const double* sp) {  // spatial.py:573-579, 672-678, 941-947
  if (type == GPY_SPATIAL_POINT) return 0.0;
  if (type == GPY_SPATIAL_GAUSS) return 5 * sp[2];
  return 1.1 * sp[2] * (1 + sp[5]);
}

double evaluation_bin_size_min(int type, const double* sp) {  // spatial.py:568-571, 667-670, 933-939
  if (type == GPY_SPATIAL_POINT) return 0.0;
  if (type == GPY_SPATIAL_GAUSS) return sp[2] / 3.0;
  return sp[2] * (1 - sp[5]) / 10.0;
}

// DiskSpatialModel._evaluate_norm_factor (spatial.py:951-969).  The reference integrates with
// scipy.integrate.quad; the integrand is smooth and pi-periodic, so a composite Gauss-Legendre rule
// converges to double precision (the oracle integrates with scipy's quad: tests/test_gpu_parity.py, case
// "disk-elongated").
double disk_norm_factor(double r0, double e) {
  double semi_minor = r0 * std::sqrt(1 - e * e);
  double A = 1 / (std::sin(r0) * std::sin(r0));
  double B = 1 / (std::sin(semi_minor) * std::sin(semi_minor));
  double C = A - B;
  auto fcn = [&](double x) {
    double cs2 = std::cos(x) * std::cos(x);
    return 1 - std::sqrt(1 - 1 / (B + C * cs2));
  };
  double integral;
  if (e == 0) {
    integral = M_PI * fcn(0.0);
  } else {
    // 8-point Gauss-Legendre on 64 panels
    static const double xg[4] = {0.1834346424956498, 0.5255324099163290, 0.7966664774136267, 0.9602898564975363};
    static const double wg[4] = {0.3626837833783620, 0.3137066458778873, 0.2223810344533745, 0.1012285362903763};
    const int panels = 64;
    double hpan = M_PI / panels, s = 0.0;
    for (int p = 0; p < panels; ++p) {
      double c = (p + 0.5) * hpan, hw = 0.5 * hpan;
      for (int k = 0; k < 4; ++k) s += wg[k] * (fcn(c - hw * xg[k]) + fcn(c + hw * xg[k]));
    }
    integral = s * 0.5 * hpan;
  }
  return 1.0 / (2 * integral);
}

// ---------------------------------------------------------------------------------------------
// runtime objects
// ---------------------------------------------------------------------------------------------
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaStream_t pool_stream = nullptr;  // non-null: stream-ordered allocation (cudaMallocAsync pool, no device sync)
  int reserve(size_t bytes, cudaStream_t stream = nullptr) {
    if (bytes <= cap) return GPY_OK;
    release();
    size_t want = std::max(bytes, (size_t)256);
    if (stream) {
      CUDA_TRY(cudaMallocAsync(&p, want, stream));
      pool_stream = stream;
    } else {
      CUDA_TRY(cudaMalloc(&p, want));
      pool_stream = nullptr;
    }
    cap = want;
    return GPY_OK;
  }
  void release() {
    if (p) {
      if (pool_stream)
        cudaFreeAsync(p, pool_stream);
      else
        cudaFree(p);
    }
    p = nullptr;
    cap = 0;
  }
};

struct Slot {  // one cached integrate_geom (+ PSF convolution) result
  DevBuf w, conv;
  Rect rect;
  int x0a = 0, pitch = 0, planes = 0, f = 0;
  std::vector<double> key;  // spatial parameter values
  bool valid = false;
  uint64_t stamp = 0;
  uint64_t call_id = 0;  // evaluate() call that last handed this slot to a kernel: never evicted within that call
};

struct Source {
  gpy_source_desc desc;
  bool initialised = false;  // MapEvaluator.exposure is not None
  bool contributes = true;
  Rect rect;                 // exposure cutout
  double cached_lon = 0.0, cached_lat = 0.0;  // evaluator.py:93 _cached_position (radians)
  std::vector<double> cached_spatial;         // _cached_parameter_values_spatial
  bool has_cached_spatial = false;
  std::vector<Slot> slots = std::vector<Slot>(kSlotsPerSource);
  int last_f = 0;
  int n_spatial_evals = 0;
  // position-dependent IRFs (gpy_dataset_set_irf_maps): this evaluator's own PSF kernel and edisp matrix, looked up
  // at the source position by MapEvaluator.update (evaluator.py:196-227)
  DevBuf p_kflip, p_planes, p_raw, p_kernel, p_jobs;
  std::vector<gpy::ConvPlane> planes;
  int psf_k = 0, max_h = 0, max_kw = 4;
  int edisp_pix = -1;    // IRF pixel (iy * nx + ix) whose matrix this evaluator uses
  int n_irf_updates = 0;
  double lookup_lon = 0.0, lookup_lat = 0.0;  // position of the last lookup (radians)
  void release_irf() {
    for (DevBuf* b : {&p_kflip, &p_planes, &p_raw, &p_kernel, &p_jobs}) b->release();
  }
};

struct IrfMaps {  // PSFMap / EDispKernelMap with a spatial grid (host copies; the maps are small)
  bool has_psf = false, has_edisp = false;
  GeomH geom{0, 0, 0.0, 0.0, 0.0, 0.0};
  int n_rad = 0;
  std::vector<double> rad_edges, rad_center;
  std::vector<float> psf;     // [nT][n_rad][ny][nx]
  std::vector<double> edisp;  // [nT][nR][ny][nx]
  DevBuf d_rad_center;
};

}  // namespace

struct gpy_context {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int64_t launches = 0;
  uint64_t stamp = 0;
  void* pinned = nullptr;
  size_t pinned_cap = 0;
  cudaEvent_t h2d_done = nullptr;
  cudaEvent_t t_begin = nullptr, t_end = nullptr;  // gpy_kernel_timing
  bool timing = false, timing_pending = false;
  double timing_ms = 0.0;
  int64_t timing_launches = 0;
  bool h2d_pending = false;
  DevBuf zeros, desc, dedup, staging, flux, partials, stat, tiles, scratch_a, scratch_b, scratch_c;
  std::vector<gpy_dataset*> tiles_for;  // dataset list the tile table was built for
  std::vector<uint64_t> tiles_ver;
  double* stat_pinned = nullptr;
  int* nwork_pinned = nullptr;  // read-back of the batched worklist length
  size_t stat_pinned_cap = 0;
  size_t ts_smem_set = 0, fold_smem_set = 0, conv_smem_set = 0, tma_smem_set[24] = {0};
  int tma_blocks_per_sm[2] = {0, 0};
  int num_sms = 0;
  bool force_generic = false;
  bool static_schedule = false;  // GPY_STATIC_SCHEDULE: never claim items dynamically
  int t64_stages = 1;            // GPY_FOLD_T64 = 1 (three CTAs per SM, single-buffered inputs) or 2 (two CTAs, double-buffered)
  bool use_t64 = true;           // GPY_FOLD_T64=0 switches the 64-pixel sub-tile kernel off (default: used whenever the launch is eligible)
  // Spatial chains (integrate_geom + PSF stencil) of the slots this evaluation has to fill: queued by get_slot, packed
  // into the descriptor blob and run as three launches for all of them before the fold (flush in evaluate_impl).
  struct PendingSpatial {
    gpy::SpatialBatchJob job;
    bool oversample;
    size_t conv_smem;
    Source* src;
    size_t slot;  // index into src->slots
  };
  DevBuf joint_dev;  // gpy_stat_sum_joint: the all-reduced statistics before they are copied to the host
  std::vector<PendingSpatial> pending;
  size_t conv_batch_smem_set = 0;
  bool batch_spatial = true;  // GPY_SPATIAL_BATCH=0: three launches per moved source instead (A/B, tests)
  int max_ctas_per_sm = 0;       // GPY_FOLD_CTAS_PER_SM (profiling): cap on resident fold CTAs per SM, 0 = occupancy limit
  bool no_desc_cache = false;  // GPY_NO_DESC_CACHE: run prep_items_kernel on every call
  std::vector<uint8_t> desc_key;  // what the descriptors in `desc` were built from (empty: none)
  bool no_dedup = false;  // GPY_NO_DEDUP: evaluate every item of a batch even when identical to vector 0's
  int fold_mode = 0;  // GPY_FOLD_KERNEL: 0 auto (bulk-copy staged kernel when eligible), 2 generic
  double log_trunc = 0.0;
  int max_smem_optin = 0;
  // ---- joint fits over ranks: one-shot all-reduce over NVLink peer memory (gpy_peer_reduce_init) ------------------
  gpy::PeerReduceParams peer{};
  bool peer_on = false;
  DevBuf peer_seq;
  // ---- steady-state replay (evaluate_replay): the step of a fit whose spatial parameters did not move -------------
  struct Replay {
    bool valid = false;
    bool disabled = false;             // GPY_NO_REPLAY
    std::vector<gpy_dataset*> datasets;
    std::vector<uint64_t> versions;
    int n_par = 0, include_background = 0;
    double* joint_out = nullptr;       // the captured step ends with the peer all-reduce into this buffer (or null)
    std::vector<int> spatial_cols;     // parameter columns whose values the cached work depends on ...
    std::vector<double> spatial_vals;  // ... and their values at capture time (bitwise)
    std::vector<int> amp_cols;         // amplitude columns: a source leaves the list when its amplitude is exactly 0
    std::vector<uint8_t> amp_zero;
    cudaGraphExec_t exec = nullptr;    // built at the first replay after a capture (a step that moves a source never pays)
    bool exec_valid = false;
    // launch parameters of the captured step
    unsigned k0_blocks = 0, k0_threads = 0;
    int k0_seg_cap = 0;
    const gpy::FluxJob* k0_jobs = nullptr;
    void (*fold_kernel)(const gpy::FoldParams) = nullptr;
    unsigned fold_grid = 0;
    size_t fold_smem = 0;
    gpy::FoldParams fold_params;
    int n_tiles = 0;
    int kernels = 0;                   // kernel launches inside the graph
    DevBuf d_params;
    double* pinned = nullptr;          // ring of parameter vectors
    std::vector<cudaEvent_t> slot_done;
    int slot = 0;
    size_t pinned_doubles = 0;
    int64_t replays = 0;
  } replay;
};

struct gpy_dataset {
  gpy_context* ctx = nullptr;
  GeomH geom;
  int nT = 0, nR = 0, nRp = 0, n_groups = 1, n_nodes = 0, psf_k = 0;
  uint64_t version = 0;
  const float *exposure = nullptr, *background = nullptr, *counts = nullptr;
  const uint8_t* mask = nullptr;
  const float* weight = nullptr;
  int compact = 0;            // C5 storage: counts uint16, mask bit-packed with rows of mask_row_bytes
  int64_t mask_row_bytes = 0;
  std::vector<uint8_t> mask_image;  // empty = all true
  DevBuf d_edges, d_nodes, d_ecenter, d_edisp, d_band, d_pdfc, d_omega, d_kflip, d_planes;
  DevBuf d_tmaps;             // 4 CUtensorMap: exposure, background, counts, mask (2-D [rows][P] tensors, box [rows][128])
  std::vector<DevBuf> retired_tmaps;  // earlier map blocks: a rewritten map gets a NEW address (descriptor caches)
  bool has_tmaps = false, has_tmaps64 = false;
  std::vector<gpy::ConvPlane> planes;
  int max_h = 0, max_kw = 4;
  bool has_psf = false, has_diag = false;
  int groups_used = 1;  // 2 when some source opts out of the edisp (diagonal response)
  std::vector<Source> sources;
  gpy_background_desc bkg{0, -1, -1, -1};
  IrfMaps irf;
  std::vector<double> h_edisp, h_diag;  // host copies of the static matrices (group tables are rebuilt from them)
  std::vector<int> group_codes;         // matrix behind every edisp group: -2 dataset edisp, -1 diagonal, >= 0 IRF pixel
  std::vector<int> pdfc_rows_cum;       // band-compressed rows up to and including group g
};

namespace {

int ensure_pinned(gpy_context* c, size_t bytes) {
  if (bytes <= c->pinned_cap) return GPY_OK;
  if (c->pinned) cudaFreeHost(c->pinned);
  c->pinned = nullptr;
  c->pinned_cap = 0;
  size_t want = std::max(bytes * 2, (size_t)1 << 16);
  CUDA_TRY(cudaMallocHost(&c->pinned, want));
  c->pinned_cap = want;
  return GPY_OK;
}

int ensure_stat_pinned(gpy_context* c, size_t n) {
  if (n <= c->stat_pinned_cap) return GPY_OK;
  if (c->stat_pinned) cudaFreeHost(c->stat_pinned);
  c->stat_pinned = nullptr;
  c->stat_pinned_cap = 0;
  size_t want = std::max(n * 2, (size_t)1024);
  CUDA_TRY(cudaMallocHost((void**)&c->stat_pinned, want * sizeof(double)));
  c->stat_pinned_cap = want;
  return GPY_OK;
}

struct Packer {  // lays out the per-call descriptor block, uploaded with one cudaMemcpyAsync
  std::vector<uint8_t> bytes;
  size_t add(const void* src, size_t n, size_t align = 16) {
    size_t off = (bytes.size() + align - 1) / align * align;
    bytes.resize(off + n);
    if (n) std::memcpy(bytes.data() + off, src, n);
    return off;
  }
};

int upload_to(DevBuf& buf, const void* src, size_t bytes, cudaStream_t s) {
  GPY_TRY(buf.reserve(bytes));
  if (bytes) CUDA_TRY(cudaMemcpyAsync(buf.p, src, bytes, cudaMemcpyHostToDevice, s));
  return GPY_OK;
}

gpy::WcsDev wcs_of(const GeomH& g) { return gpy::WcsDev{g.crval_lon, g.binsz, g.crpix_x, g.crpix_y, g.crval_lat, g.proj}; }

// ---- K1 + K2 into a slot ---------------------------------------------------------------------
// SpatialModel.integrate_geom on the evaluator geometry `rect` of dataset geometry `pg`, then
// WcsNDMap.convolve.  `omega` is the parent solid-angle image (device).
// rows_out[2] (optional): the rows [r0, r1) of the cutout outside of which the image is exactly zero.
int compute_spatial(gpy_context* c, const GeomH& pg, const double* omega, const Rect& rect, int type,
                    const double* sp, float* w_out, int pitch, int xpad, int* f_out, int* rows_out = nullptr,
                    gpy::SpatialJob* job_out = nullptr,  // job_out: describe the work instead of launching it
                    int model_frame = -1) {              // gpy_frame of sp (lon_0, lat_0, phi); -1: the map's
  GeomH eg = sub_geom(pg, rect);
  gpy::SpatialJob j;
  std::memset(&j, 0, sizeof(j));
  j.type = type;
  j.ecx = rect.cx;
  j.ecy = rect.cy;
  j.pitch = pitch;
  j.xpad = xpad;
  j.ex0 = rect.x0;
  j.ey0 = rect.y0;
  j.parent_nx = pg.nx;
  j.wcs_e = wcs_of(eg);
  j.solid_angle = omega;
  j.out = w_out;
  j.lon0 = sp[0] * kDegH;
  j.lat0 = sp[1] * kDegH;
  // the position in the map's frame (pixel arithmetic: skycoord_to_pixel) and, for a model in another frame, the
  // rotation that takes map coordinates to the model's
  double mlon = sp[0], mlat = sp[1];
  if (model_frame >= 0 && model_frame != pg.frame) {
    frame_convert(model_frame, pg.frame, sp[0], sp[1], mlon, mlat);
    j.rotate = frame_matrix(pg.frame, model_frame, j.rot) ? 1 : 0;
  }
  int f = 1;
  if (type == GPY_SPATIAL_POINT) {
    world_to_pix(eg, mlon, mlat, j.px0, j.py0);
  } else {
    const double pix_scale = pg.binsz;
    const double radius = evaluation_radius(type, sp);
    bool have_f = false;
    Rect inner;
    // try: wcs_geom.cutout(position, 2 * max(r_eval, pix_scale)) (spatial.py:255-264)
    double width = 2 * (radius > pix_scale || std::isnan(radius) ? radius : pix_scale);  // np.maximum
    CutStatus cs = cutout_rect(eg, mlon, mlat, width, false, inner);
    if (cs != CUT_OK) {
      f = 1;
      have_f = true;
    }
    if (!have_f) {
      double res = evaluation_bin_size_min(type, sp);
      if (res > 0) {
        double ratio = std::ceil(pix_scale / res);
        f = ratio > (double)kMaxOversampling ? kMaxOversampling : (int)ratio;
        if (f > kMaxOversampling) f = kMaxOversampling;
      } else {
        f = kMaxOversampling;
      }
    }
    double e = sp[3];
    double phi = py_mod(sp[4] * kDegH - 0.0, M_PI - 0.0) + 0.0;  // wrap_at(phi, 0, 180 deg)
    j.ecc = e;
    j.phi = phi;
    j.major = sp[2] * kDegH;
    if (type == GPY_SPATIAL_GAUSS) {
      double sigma = sp[2] * kDegH;
      if (e == 0) {
        j.a = 1.0 - std::cos(sigma);
        j.norm = 1 / (4 * M_PI * j.a * (1.0 - std::exp(-1.0 / j.a)));
      } else {
        double minor = sigma * std::sqrt(1 - e * e);
        j.norm = 1 / (2 * M_PI * sigma * minor);
      }
    } else {
      j.edge_width = sp[5];
      j.norm = disk_norm_factor(sp[2] * kDegH, e);
    }
    if (f > 1) {
      j.ix0 = inner.x0;
      j.iy0 = inner.y0;
      j.icx = inner.cx;
      j.icy = inner.cy;
      GeomH cg = sub_geom(eg, inner);
      // WcsGeom.upsample + get_resampled_wcs: crpix' = (crpix - 0.5) * f + 0.5, cdelt' = cdelt / f
      j.wcs_up = wcs_of(cg);  // (projection and reference point)
      j.wcs_up.binsz = cg.binsz / f;
      j.wcs_up.crpix_x = (cg.crpix_x - 0.5) * f + 0.5;
      j.wcs_up.crpix_y = (cg.crpix_y - 0.5) * f + 0.5;
    }
  }
  j.f = f;
  if (f_out) *f_out = f;
  if (rows_out) {
    rows_out[0] = 0;
    rows_out[1] = rect.cy;
    if (type == GPY_SPATIAL_POINT && std::isfinite(j.py0)) {  // weights (1 - |y - py0|)+: rows within one pixel of py0
      rows_out[0] = std::max(0, (int)std::ceil(j.py0 - 1.0));
      rows_out[1] = std::min(rect.cy, (int)std::floor(j.py0 + 1.0) + 1);
    } else if (type != GPY_SPATIAL_POINT && f > 1) {          // zeros outside the oversampled inner cutout
      rows_out[0] = std::max(0, j.iy0);
      rows_out[1] = std::min(rect.cy, j.iy0 + j.icy);
    }
    if (rows_out[1] < rows_out[0]) rows_out[1] = rows_out[0];
  }
  if (job_out) {
    *job_out = j;
    return GPY_OK;
  }
  dim3 grid((pitch + 127) / 128, rect.cy);
  gpy::spatial_fill_kernel<<<grid, 128, 0, c->stream>>>(j);
  c->launches++;
  if (type != GPY_SPATIAL_POINT && f > 1) {
    long long warps = (long long)j.icx * j.icy;
    int blocks = (int)((warps * 32 + 255) / 256);
    gpy::spatial_oversample_kernel<<<blocks, 256, 0, c->stream>>>(j);
    c->launches++;
  }
  CUDA_TRY(cudaGetLastError());
  return GPY_OK;
}

size_t conv_smem_bytes(int h, int kw) {
  return ((size_t)(gpy::kConvTY + 2 * h) * (gpy::kConvTX + kw) + (size_t)(2 * h + 1) * kw) * sizeof(float);
}

int launch_conv(gpy_context* c, const float* image, int cy, int cx, int pitch, int xpad, const float* kflip,
                const gpy::ConvPlane* planes_dev, int n_planes, int max_h, int max_kw, float* out, int row0 = 0,
                int row1 = 1 << 30) {
  size_t smem = conv_smem_bytes(max_h, max_kw);
  if (smem > (size_t)c->max_smem_optin)
    return fail(GPY_ERR_UNSUPPORTED, "PSF kernel too large for the shared-memory stencil (half width %d)", max_h);
  if (smem > c->conv_smem_set) {
    CUDA_TRY(cudaFuncSetAttribute(gpy::psf_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    c->conv_smem_set = smem;
  }
  dim3 grid((pitch + gpy::kConvTX - 1) / gpy::kConvTX, (cy + gpy::kConvTY - 1) / gpy::kConvTY, n_planes);
  gpy::psf_conv_kernel<<<grid, gpy::kConvThreads, smem, c->stream>>>(image, cy, cx, pitch, xpad, kflip, planes_dev,
                                                                     out, (size_t)cy * pitch, row0, std::min(row1, cy));
  c->launches++;
  CUDA_TRY(cudaGetLastError());
  return GPY_OK;
}

// Flip, crop to the non-zero support and pad each kernel plane (maps/wcs/ndmap.py:961: float32 planes).
void prepare_planes(const float* kernel, int n_planes, int K, std::vector<float>& kflip,
                    std::vector<gpy::ConvPlane>& planes, int& max_h, int& max_kw) {
  const int hK = (K - 1) / 2;
  kflip.clear();
  planes.resize(n_planes);
  max_h = 0;
  max_kw = 4;
  for (int t = 0; t < n_planes; ++t) {
    const float* kp = kernel + (size_t)t * K * K;
    int h = 0;
    for (int i = 0; i < K; ++i)
      for (int jx = 0; jx < K; ++jx)
        if (kp[(size_t)i * K + jx] != 0.0f) h = std::max(h, std::max(std::abs(i - hK), std::abs(jx - hK)));
    int kw = (2 * h + 1 + 3) / 4 * 4;
    gpy::ConvPlane pl{h, kw, (int)kflip.size(), 0};
    kflip.resize(kflip.size() + (size_t)(2 * h + 1) * kw, 0.0f);
    for (int i = 0; i <= 2 * h; ++i)
      for (int jx = 0; jx <= 2 * h; ++jx)
        kflip[pl.koff + (size_t)i * kw + jx] = kp[(size_t)(hK + h - i) * K + (hK + h - jx)];
    planes[t] = pl;
    max_h = std::max(max_h, h);
    max_kw = std::max(max_kw, kw);
  }
}

// 2-D tensor maps for the bulk-tensor (TMA) loads of the fold kernel: tensor [rows][P] (P innermost), box [rows][128].
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// esz: element size 4 (float32), 2 (uint16) or 1 (uint8); `inner` elements per row, rows `row_bytes` apart, box of
// `box_inner` elements x all rows.
bool encode_map(CUtensorMap* out, const void* base, int esz, size_t inner, size_t row_bytes, int rows, int box_inner) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn || !base || rows > 256 || row_bytes % 16 != 0) return false;
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)row_bytes};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)rows};
  cuuint32_t estr[2] = {1, 1};
  const CUtensorMapDataType dt = esz == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                 : esz == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT16 : CU_TENSOR_MAP_DATA_TYPE_UINT8;
  return fn(out, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// (Re)build the dataset's tensor maps; leaves has_tmaps false when the shape is not eligible (the kernel then uses
// 1-D bulk copies) or the driver entry point is missing.
int build_tmaps(gpy_dataset* ds) {
  ds->has_tmaps = false;
  const size_t P = (size_t)ds->geom.nx * ds->geom.ny;
  if (P % 16 != 0 || ds->geom.nx % 4 != 0 || P < (size_t)gpy::kFoldTPX) return GPY_OK;
  alignas(64) CUtensorMap maps[8];  // [0, 4): 128-pixel boxes, [4, 8): 64-pixel boxes (fold_cash_t64_kernel)
  std::memset(maps, 0, sizeof(maps));
  const int TPX = gpy::kFoldTPX;
  if (!encode_map(&maps[0], ds->exposure, 4, P, P * 4, ds->nT, TPX)) return GPY_OK;
  if (ds->background && !encode_map(&maps[1], ds->background, 4, P, P * 4, ds->nR, TPX)) return GPY_OK;
  if (!ds->compact) {
    if (ds->counts && !encode_map(&maps[2], ds->counts, 4, P, P * 4, ds->nR, TPX)) return GPY_OK;
    if (ds->mask && !encode_map(&maps[3], ds->mask, 1, P, P, ds->nR, TPX)) return GPY_OK;
  } else {  // uint16 counts; mask bits: 16 bytes per tile and row
    if (ds->counts && !encode_map(&maps[2], ds->counts, 2, P, P * 2, ds->nR, TPX)) return GPY_OK;
    if (ds->mask && !encode_map(&maps[3], ds->mask, 1, (size_t)ds->mask_row_bytes, (size_t)ds->mask_row_bytes, ds->nR, TPX / 8))
      return GPY_OK;
  }
  ds->has_tmaps64 = encode_map(&maps[4], ds->exposure, 4, P, P * 4, ds->nT, TPX / 2) &&
                    (!ds->background || encode_map(&maps[5], ds->background, 4, P, P * 4, ds->nR, TPX / 2));
  if (ds->has_tmaps64 && !ds->compact)
    ds->has_tmaps64 = (!ds->counts || encode_map(&maps[6], ds->counts, 4, P, P * 4, ds->nR, TPX / 2)) &&
                      (!ds->mask || encode_map(&maps[7], ds->mask, 1, P, P, ds->nR, TPX / 2));
  else if (ds->has_tmaps64)
    ds->has_tmaps64 = (!ds->counts || encode_map(&maps[6], ds->counts, 2, P, P * 2, ds->nR, TPX / 2)) &&
                      (!ds->mask || encode_map(&maps[7], ds->mask, 1, (size_t)ds->mask_row_bytes, (size_t)ds->mask_row_bytes,
                                               ds->nR, TPX / 8));
  if (ds->d_tmaps.p) {  // never rewrite a tensor map in place: the GPU may hold the old one in its descriptor cache
    ds->retired_tmaps.push_back(ds->d_tmaps);
    ds->d_tmaps = DevBuf();
  }
  GPY_TRY(ds->d_tmaps.reserve(sizeof(maps)));
  CUDA_TRY(cudaMemcpyAsync(ds->d_tmaps.p, maps, sizeof(maps), cudaMemcpyHostToDevice, ds->ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ds->ctx->stream));
  ds->has_tmaps = true;
  return GPY_OK;
}

// Edisp groups of a dataset: matrices [G][nT][nRp] with the reco axis zero padded to a multiple of 8, the non-zero
// band of every (group, reco chunk) and the band-compressed copy the staged fold kernel keeps in shared memory.
// `mats[g]` is a HOST [nT][nR] matrix.
int build_edisp_tables(gpy_dataset* ds, const std::vector<const double*>& mats) {
  const int G = (int)mats.size(), nT = ds->nT, nR = ds->nR, nRp = ds->nRp;
  const int nchunks = nRp / 8;
  std::vector<double> ed((size_t)G * nT * nRp, 0.0);
  std::vector<int> band((size_t)G * nchunks * 4, 0);
  std::vector<double> pdfc;
  ds->pdfc_rows_cum.assign(G, 0);
  for (int gi = 0; gi < G; ++gi) {
    const double* src = mats[gi];
    auto at = [&](int t, int r) { return src[(size_t)t * nR + r]; };
    for (int t = 0; t < nT; ++t)
      for (int r = 0; r < nR; ++r) ed[((size_t)gi * nT + t) * nRp + r] = at(t, r);
    for (int ch = 0; ch < nchunks; ++ch) {
      int lo = nT, hi = 0;
      for (int t = 0; t < nT; ++t)
        for (int r = 8 * ch; r < std::min(8 * ch + 8, nR); ++r) {
          double v = at(t, r);
          if (v != 0.0) {  // NaN counts as non-zero
            lo = std::min(lo, t);
            hi = std::max(hi, t + 1);
          }
        }
      if (lo > hi) lo = hi = 0;
      int* b4 = &band[((size_t)gi * nchunks + ch) * 4];
      b4[0] = lo;
      b4[1] = hi;
      b4[2] = (int)(pdfc.size() / 8);  // first row of this chunk in the band-compressed matrix
      for (int t = lo; t < hi; ++t)
        for (int j = 0; j < 8; ++j) {
          const int r = 8 * ch + j;
          pdfc.push_back(r < nR ? at(t, r) : 0.0);
        }
    }
    ds->pdfc_rows_cum[gi] = (int)(pdfc.size() / 8);
  }
  if (pdfc.empty()) pdfc.resize(8, 0.0);
  cudaStream_t s = ds->ctx->stream;
  // new buffers: kernels of earlier calls may still read the old tables (stream order frees them afterwards)
  DevBuf n_pdfc, n_ed, n_band;
  GPY_TRY(n_pdfc.reserve(pdfc.size() * sizeof(double), s));
  GPY_TRY(n_ed.reserve(ed.size() * sizeof(double), s));
  GPY_TRY(n_band.reserve(band.size() * sizeof(int), s));
  CUDA_TRY(cudaMemcpyAsync(n_pdfc.p, pdfc.data(), pdfc.size() * sizeof(double), cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaMemcpyAsync(n_ed.p, ed.data(), ed.size() * sizeof(double), cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaMemcpyAsync(n_band.p, band.data(), band.size() * sizeof(int), cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaStreamSynchronize(s));  // the staging vectors go out of scope
  ds->d_pdfc.release();
  ds->d_edisp.release();
  ds->d_band.release();
  ds->d_pdfc = n_pdfc;
  ds->d_edisp = n_ed;
  ds->d_band = n_band;
  ds->n_groups = G;
  return GPY_OK;
}

// ---- IRF lookup at a source position -------------------------------------------------------------
// Fractional pixel of the IRF grid, clamped to the grid (PSFMap._get_nearest_valid_position keeps the lookup inside).
void irf_pixel(const IrfMaps& m, double lon, double lat, double& px, double& py) {
  world_to_pix(m.geom, lon, lat, px, py);
  px = std::min(std::max(px, 0.0), m.geom.nx - 1.0);
  py = std::min(std::max(py, 0.0), m.geom.ny - 1.0);
}

// PSF.containment_radius (irf/psf/core.py:39-87) on the profile of one position: containment = cumulative sum of
// psf * bin width * 2 pi rad (IRF.cumsum, irf/core.py:358-390) at the upper bin edges, linear interpolation with 0
// outside the nodes, argmin of |containment - fraction| on the 20 x upsampled rad axis.
void containment_radius_h(const IrfMaps& m, const std::vector<double>& profile, int nT, double fraction,
                          std::vector<double>& out) {
  const int n = m.n_rad, factor = 20, n_up = n * factor;
  const std::vector<double>& e = m.rad_edges;
  out.assign(nT, 0.0);
  std::vector<double> cum(n);
  const double step = (e[n] - e[0]) / n_up;
  for (int t = 0; t < nT; ++t) {
    double acc = 0.0;
    for (int k = 0; k < n; ++k) {
      acc += profile[(size_t)t * n + k] * ((e[k + 1] - e[k]) * kDegH) * (2 * M_PI * m.rad_center[k] * kDegH);
      cum[k] = acc;
    }
    double best = std::numeric_limits<double>::infinity(), best_rad = 0.0;
    int seg = 0;
    for (int i = 0; i < n_up; ++i) {
      const double lo = i == 0 ? e[0] : e[0] + i * step, hi = i + 1 == n_up ? e[n] : e[0] + (i + 1) * step;
      const double rad = 0.5 * (lo + hi);
      double cont;  // np.interp(rad, nodes = edges[1:], cum, left=0, right=0)
      if (rad < e[1] || rad > e[n]) {
        cont = 0.0;
      } else {
        while (seg + 2 < n + 1 && rad > e[seg + 2]) ++seg;  // nodes[seg] = e[seg + 1] <= rad <= nodes[seg + 1]
        if (seg + 1 >= n) {
          cont = cum[n - 1];
        } else {
          const double x0 = e[seg + 1], x1 = e[seg + 2];
          cont = cum[seg] + (cum[seg + 1] - cum[seg]) * ((rad - x0) / (x1 - x0));
        }
      }
      const double dist = std::fabs(cont - fraction);
      if (dist < best) {  // np.argmin: first minimum
        best = dist;
        best_rad = rad;
      }
    }
    out[t] = best_rad;
  }
}

// PSFMap.get_psf_kernel(position, geom, containment=0.999) (irf/psf/map.py:249-337) for one evaluator: profile and
// containment radii on the host (a few thousand numbers), the oversampled kernel images, their normalisation and the
// stencil layout on the device (psf_extract_kernel).
int extract_psf(gpy_dataset* ds, Source& src, double lon, double lat) {
  gpy_context* c = ds->ctx;
  const IrfMaps& m = ds->irf;
  const int nT = ds->nT, n = m.n_rad, inx = m.geom.nx, iny = m.geom.ny;
  double px, py;
  irf_pixel(m, lon, lat, px, py);
  const int x0 = (int)std::min(std::floor(px), (double)std::max(inx - 2, 0));
  const int y0 = (int)std::min(std::floor(py), (double)std::max(iny - 2, 0));
  const int x1 = std::min(x0 + 1, inx - 1), y1 = std::min(y0 + 1, iny - 1);
  const double tx = px - x0, ty = py - y0;
  std::vector<double> profile((size_t)nT * n);
  const size_t plane = (size_t)inx * iny;
  for (int t = 0; t < nT; ++t)
    for (int k = 0; k < n; ++k) {
      const float* q = m.psf.data() + ((size_t)t * n + k) * plane;
      profile[(size_t)t * n + k] = (1 - ty) * ((1 - tx) * (double)q[(size_t)y0 * inx + x0] + tx * (double)q[(size_t)y0 * inx + x1]) +
                                   ty * ((1 - tx) * (double)q[(size_t)y1 * inx + x0] + tx * (double)q[(size_t)y1 * inx + x1]);
    }
  std::vector<double> r999, r68;
  containment_radius_h(m, profile, nT, 0.999, r999);
  containment_radius_h(m, profile, nT, 0.68, r68);
  const double binsz = ds->geom.binsz;
  const double max_radius = *std::max_element(r999.begin(), r999.end());
  const int K = (int)round_up_to_odd(2 * max_radius / binsz);
  std::vector<gpy::PsfPlaneJob> jobs(nT);
  std::vector<int> raw_off(nT);
  src.planes.assign(nT, gpy::ConvPlane{0, 4, 0, 0});
  int koff = 0, roff = 0;
  src.max_h = 0;
  src.max_kw = 4;
  for (int t = 0; t < nT; ++t) {
    const double base = 2 * r68[t] / binsz;
    // np.minimum(int(np.ceil(precision_factor / base_factor)), PSF_MAX_OVERSAMPLING); base 0 -> ceil(inf) raises in
    // the reference (OverflowError): refuse it here as well
    if (!(base > 0)) return fail(GPY_ERR_INVALID, "PSF map: zero 68 %% containment radius in true-energy bin %d", t);
    const int f = (int)std::min(std::ceil(12.0 / base), 4.0);
    const int nt = std::min((int)round_up_to_odd(2 * r999[t] / binsz), K);
    const int h = (nt - 1) / 2, kw = (nt + 3) / 4 * 4;
    jobs[t] = gpy::PsfPlaneJob{nt, f, koff, kw};
    src.planes[t] = gpy::ConvPlane{h, kw, koff, 0};
    raw_off[t] = roff;
    koff += nt * kw;
    roff += nt * nt;
    src.max_h = std::max(src.max_h, h);
    src.max_kw = std::max(src.max_kw, kw);
  }
  src.psf_k = K;
  cudaStream_t st = c->stream;
  // fresh buffers every time: slots convolved with the previous kernel may still be read by queued launches
  for (DevBuf* b : {&src.p_kflip, &src.p_planes, &src.p_raw, &src.p_kernel, &src.p_jobs}) {
    if (b->p) CUDA_TRY(cudaStreamSynchronize(st));
    b->release();
  }
  const size_t job_bytes = jobs.size() * sizeof(gpy::PsfPlaneJob), off_bytes = raw_off.size() * sizeof(int),
               prof_bytes = profile.size() * sizeof(double);
  const size_t o_off = (job_bytes + 15) / 16 * 16, o_prof = (o_off + off_bytes + 15) / 16 * 16;
  std::vector<uint8_t> blob(o_prof + prof_bytes);
  std::memcpy(blob.data(), jobs.data(), job_bytes);
  std::memcpy(blob.data() + o_off, raw_off.data(), off_bytes);
  std::memcpy(blob.data() + o_prof, profile.data(), prof_bytes);
  GPY_TRY(upload_to(src.p_jobs, blob.data(), blob.size(), st));
  GPY_TRY(upload_to(src.p_planes, src.planes.data(), src.planes.size() * sizeof(gpy::ConvPlane), st));
  GPY_TRY(src.p_kflip.reserve((size_t)std::max(koff, 1) * sizeof(float)));
  GPY_TRY(src.p_raw.reserve((size_t)std::max(roff, 1) * sizeof(double)));
  GPY_TRY(src.p_kernel.reserve((size_t)nT * K * K * sizeof(float)));
  CUDA_TRY(cudaMemsetAsync(src.p_kflip.p, 0, (size_t)std::max(koff, 1) * sizeof(float), st));
  CUDA_TRY(cudaMemsetAsync(src.p_kernel.p, 0, (size_t)nT * K * K * sizeof(float), st));
  gpy::PsfExtractParams q;
  q.jobs = (const gpy::PsfPlaneJob*)src.p_jobs.p;
  q.raw_off = (const int*)((const uint8_t*)src.p_jobs.p + o_off);
  q.profile = (const double*)((const uint8_t*)src.p_jobs.p + o_prof);
  q.rad_center = (const double*)m.d_rad_center.p;
  q.n_rad = n;
  q.n_planes = nT;
  q.binsz = binsz;
  q.proj = ds->geom.proj;
  q.pad_ = 0;
  q.raw = (double*)src.p_raw.p;
  q.kflip = (float*)src.p_kflip.p;
  q.kernel = (float*)src.p_kernel.p;
  q.K = K;
  gpy::psf_extract_kernel<<<nT, 256, (2 * n + 256) * sizeof(double), st>>>(q);
  c->launches++;
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaStreamSynchronize(st));  // the host blob goes out of scope
  src.n_irf_updates++;
  return GPY_OK;
}

// EDispKernelMap.get_edisp_kernel(position) (irf/edisp/map.py:369-401 -> to_region_nd_map with a point region:
// interp_by_coord(method="nearest")): the matrix of the nearest IRF pixel.
int edisp_pixel_at(const IrfMaps& m, double lon, double lat) {
  double px, py;
  irf_pixel(m, lon, lat, px, py);
  const int i = std::min(std::max((int)std::floor(px + 0.5), 0), m.geom.nx - 1);
  const int j = std::min(std::max((int)std::floor(py + 0.5), 0), m.geom.ny - 1);
  return j * m.geom.nx + i;
}

// MapEvaluator.update geometry (evaluator.py:174-243) for spatial parameters `sp`.
int evaluator_update(const gpy_dataset* ds, const Source& src, const double* sp, Rect& rect, bool& contributes) {
  const int type = src.desc.spatial_type;
  const int psf_k = ds->irf.has_psf ? src.psf_k : ds->psf_k;  // this evaluator's own kernel when the PSF map has a grid
  const double psf_width = ds->has_psf ? ds->geom.binsz * psf_k : 0.0;  // np.max(kernel geom width)
  const double radius = evaluation_radius(type, sp);
  const double cutout_width = psf_width + 2 * (radius + kCutoutMargin);  // evaluator.py:169-172
  // SkyModel.contributes(mask, margin=psf_width) (cube.py:246-286)
  Rect mrect;
  double mlon, mlat;  // model.position in the map's frame
  frame_convert(src.desc.frame, ds->geom.frame, sp[0], sp[1], mlon, mlat);
  CutStatus cs = cutout_rect(ds->geom, mlon, mlat, 2 * radius + kCutoutMargin + psf_width, false, mrect);
  contributes = false;
  if (cs == CUT_OK) {
    if (ds->mask_image.empty()) {
      contributes = true;
    } else {
      for (int y = mrect.y0; y < mrect.y0 + mrect.cy && !contributes; ++y) {
        const uint8_t* row = ds->mask_image.data() + (size_t)y * ds->geom.nx;
        for (int x = mrect.x0; x < mrect.x0 + mrect.cx; ++x)
          if (row[x]) {
            contributes = true;
            break;
          }
      }
    }
  }
  if (!contributes) return GPY_OK;
  cs = cutout_rect(ds->geom, mlon, mlat, cutout_width, true, rect);
  if (cs == CUT_VALUE_ERROR)
    return fail(GPY_ERR_INVALID, "non-finite source position / size in MapEvaluator.update (Cutout2D raises ValueError)");
  if (cs == CUT_NO_OVERLAP)
    return fail(GPY_ERR_INVALID, "source cutout does not overlap the dataset geometry (NoOverlapError)");
  return GPY_OK;
}

// Find or build the slot holding integrate_geom (x) PSF for (sp, rect).  Slots handed out during the current
// evaluate() call are pinned (their buffers are referenced by kernels not yet run); when every slot is pinned
// the list grows, so a batch may hold any number of distinct spatial parameter tuples.
int get_slot(gpy_dataset* ds, Source& src, const double* sp, const Rect& rect, uint64_t call_id, Slot** out) {
  gpy_context* c = ds->ctx;
  const int nsp = n_spatial_params(src.desc.spatial_type);
  Slot* victim = nullptr;
  for (auto& s : src.slots) {
    if (s.valid && s.rect == rect && (int)s.key.size() == nsp &&
        std::memcmp(s.key.data(), sp, nsp * sizeof(double)) == 0) {
      s.stamp = ++c->stamp;
      s.call_id = call_id;
      *out = &s;
      return GPY_OK;
    }
    if (s.valid && s.call_id == call_id) continue;  // pinned
    if (!victim || (!s.valid && victim->valid) || (s.valid == victim->valid && s.stamp < victim->stamp)) victim = &s;
  }
  if (!victim) {
    src.slots.emplace_back();
    victim = &src.slots.back();
  }
  Slot& s = *victim;
  s.valid = false;
  s.rect = rect;
  s.x0a = rect.x0 & ~3;
  s.pitch = (rect.x0 + rect.cx - s.x0a + 3) / 4 * 4;
  const bool conv = ds->has_psf && src.desc.apply_psf;
  s.planes = conv ? ds->nT : 1;
  const size_t img = (size_t)rect.cy * s.pitch;
  // A cutout changes its size by a pixel or two as the source moves (odd npix, ceil rules, trimming at the map
  // edge): allocate with headroom, a pool allocation that has to grow costs milliseconds in the middle of a fit.
  const size_t img_cap = (size_t)(rect.cy + 4) * (s.pitch + 8);
  if (img * sizeof(float) > s.w.cap) GPY_TRY(s.w.reserve(img_cap * sizeof(float), c->stream));
  int wrows[2] = {0, rect.cy};
  if (conv && img * ds->nT * sizeof(float) > s.conv.cap)
    GPY_TRY(s.conv.reserve(img_cap * ds->nT * sizeof(float), c->stream));
  if (c->batch_spatial) {  // queued: evaluate_impl runs the chains of all the slots of this call in three launches
    gpy_context::PendingSpatial ps;
    std::memset(&ps.job, 0, sizeof(ps.job));
    GPY_TRY(compute_spatial(c, ds->geom, (const double*)ds->d_omega.p, rect, src.desc.spatial_type, sp,
                            (float*)s.w.p, s.pitch, rect.x0 - s.x0a, &s.f, wrows, &ps.job.j, src.desc.frame));
    ps.oversample = src.desc.spatial_type != GPY_SPATIAL_POINT && s.f > 1;
    ps.job.fill_gx = (s.pitch + 127) / 128;
    ps.conv_smem = 0;
    if (conv) {
      const int max_h = ds->irf.has_psf ? src.max_h : ds->max_h, max_kw = ds->irf.has_psf ? src.max_kw : ds->max_kw;
      ps.conv_smem = conv_smem_bytes(max_h, max_kw);
      if (ps.conv_smem > (size_t)c->max_smem_optin)
        return fail(GPY_ERR_UNSUPPORTED, "PSF kernel too large for the shared-memory stencil (half width %d)", max_h);
      ps.job.kflip = (const float*)(ds->irf.has_psf ? src.p_kflip.p : ds->d_kflip.p);
      ps.job.planes = (const gpy::ConvPlane*)(ds->irf.has_psf ? src.p_planes.p : ds->d_planes.p);
      ps.job.conv_out = (float*)s.conv.p;
      ps.job.n_planes = ds->nT;
      ps.job.conv_gx = (s.pitch + gpy::kConvTX - 1) / gpy::kConvTX;
      ps.job.conv_gy = (rect.cy + gpy::kConvTY - 1) / gpy::kConvTY;
      ps.job.row0 = wrows[0];
      ps.job.row1 = std::min(wrows[1], rect.cy);
    }
    ps.src = &src;
    ps.slot = (size_t)(&s - src.slots.data());
    c->pending.push_back(ps);
  } else {
  GPY_TRY(compute_spatial(c, ds->geom, (const double*)ds->d_omega.p, rect, src.desc.spatial_type, sp,
                          (float*)s.w.p, s.pitch, rect.x0 - s.x0a, &s.f, wrows, nullptr, src.desc.frame));
  if (conv) {
    if (ds->irf.has_psf)
      GPY_TRY(launch_conv(c, (const float*)s.w.p, rect.cy, rect.cx, s.pitch, rect.x0 - s.x0a,
                          (const float*)src.p_kflip.p, (const gpy::ConvPlane*)src.p_planes.p, ds->nT, src.max_h,
                          src.max_kw, (float*)s.conv.p, wrows[0], wrows[1]));
    else
      GPY_TRY(launch_conv(c, (const float*)s.w.p, rect.cy, rect.cx, s.pitch, rect.x0 - s.x0a,
                          (const float*)ds->d_kflip.p, (const gpy::ConvPlane*)ds->d_planes.p, ds->nT, ds->max_h,
                          ds->max_kw, (float*)s.conv.p, wrows[0], wrows[1]));
  }
  }
  s.key.assign(sp, sp + nsp);
  s.valid = true;
  s.stamp = ++c->stamp;
  s.call_id = call_id;
  src.last_f = s.f;
  src.n_spatial_evals++;
  *out = &s;
  return GPY_OK;
}

struct EvalRequest {
  gpy_dataset* const* datasets;
  int n_datasets;
  const double* params;
  int n_vec, n_par;
  const uint8_t* source_enable;  // single-dataset npred only
  int include_background;
  double* npred_out;  // device
  double* joint_out = nullptr;  // device [n_vec]: all-reduce of the statistic over the ranks (gpy_stat_sum_joint_device)
};

int check_par(int idx, int n_par, const char* what) {
  if (idx < 0 || idx >= n_par) return fail(GPY_ERR_INVALID, "parameter index %d of %s outside [0, %d)", idx, what, n_par);
  return GPY_OK;
}

struct FluxKey {  // (dataset, source) + the bit patterns of the spectral parameters
  uint64_t source;
  uint64_t p[5];
  bool operator<(const FluxKey& o) const { return std::memcmp(this, &o, sizeof(FluxKey)) < 0; }
};

int timing_collect(gpy_context* c) {  // adds the bracketed kernel of the previous evaluation to the totals
  if (!c->timing_pending) return GPY_OK;
  CUDA_TRY(cudaEventSynchronize(c->t_end));
  float ms = 0.f;
  CUDA_TRY(cudaEventElapsedTime(&ms, c->t_begin, c->t_end));
  c->timing_ms += ms;
  c->timing_launches++;
  c->timing_pending = false;
  return GPY_OK;
}

// Build descriptors, refresh spatial caches, launch K0 + fused fold/Cash + reduction.
// Leaves stat[b * D + d] in c->stat (device).
constexpr int kReplaySlots = 32;

// Steady-state step of a fit: same datasets, one parameter vector, no spatial parameter moved and no amplitude
// crossing zero since the captured call.  Then nothing the host packs changes but the spectral / background values,
// which K0 reads from the parameter vector itself (FluxJob::pidx): the step is one small H2D copy plus the replay of a
// CUDA graph holding K0, the work-counter reset, the fold kernel and the reduction.  Returns true when it ran.
int evaluate_replay(gpy_context* c, const EvalRequest& rq, bool* done) {
  *done = false;
  gpy_context::Replay& R = c->replay;
  if (!R.valid || R.disabled || c->timing || rq.n_vec != 1 || rq.npred_out || rq.source_enable ||
      rq.n_datasets != (int)R.datasets.size() || rq.n_par != R.n_par || rq.include_background != R.include_background ||
      rq.joint_out != R.joint_out)
    return GPY_OK;
  for (int d = 0; d < rq.n_datasets; ++d)
    if (rq.datasets[d] != R.datasets[d] || rq.datasets[d]->version != R.versions[d]) return GPY_OK;
  for (size_t i = 0; i < R.spatial_cols.size(); ++i)
    if (std::memcmp(&rq.params[R.spatial_cols[i]], &R.spatial_vals[i], sizeof(double)) != 0) return GPY_OK;
  for (size_t i = 0; i < R.amp_cols.size(); ++i)
    if ((rq.params[R.amp_cols[i]] == 0.0) != (R.amp_zero[i] != 0)) return GPY_OK;
  if (!R.exec_valid) {  // second step in a row with the same spatial parameters: worth a graph
    if (R.exec) {
      cudaGraphExecDestroy(R.exec);
      R.exec = nullptr;
    }
    cudaGraph_t graph = nullptr;
    const int D = rq.n_datasets;
    CUDA_TRY(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
    R.kernels = 0;
    if (R.k0_blocks) {
      gpy::spectral_flux_kernel<<<R.k0_blocks, R.k0_threads, (size_t)R.k0_seg_cap * sizeof(double), c->stream>>>(
          R.k0_jobs, R.k0_seg_cap, (const double*)R.d_params.p, R.fold_params.counter);  // (also zeroes the work counter)
      R.kernels++;
    } else if (R.fold_params.counter) {
      cudaMemsetAsync(R.fold_params.counter, 0, sizeof(unsigned), c->stream);
    }
    R.fold_kernel<<<R.fold_grid, gpy::kFoldThreads, R.fold_smem, c->stream>>>(R.fold_params);
    gpy::reduce_stat_kernel<<<dim3(1, D), gpy::kReduceThreads, 0, c->stream>>>((const double*)c->partials.p, R.fold_params.ds, R.n_tiles, D, 1,
                                                              nullptr, (double*)c->stat.p);
    R.kernels += 2;
    if (R.joint_out) {
      gpy::PeerReduceParams q = c->peer;
      q.n_datasets = D;
      gpy::peer_allreduce_kernel<<<1, 256, 0, c->stream>>>(q, (const double*)c->stat.p, 1, R.joint_out);
      R.kernels++;
    }
    cudaError_t ce = cudaStreamEndCapture(c->stream, &graph);
    if (ce != cudaSuccess || !graph) {
      (void)cudaGetLastError();
      R.valid = false;  // capture not possible here (e.g. the caller is itself capturing): ordinary path
      return GPY_OK;
    }
    ce = cudaGraphInstantiate(&R.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ce != cudaSuccess) {
      (void)cudaGetLastError();
      R.exec = nullptr;
      R.valid = false;
      return GPY_OK;
    }
    R.exec_valid = true;
  }
  const int slot = R.slot;
  R.slot = (R.slot + 1) % kReplaySlots;
  CUDA_TRY(cudaEventSynchronize(R.slot_done[slot]));  // the copy that last used this slot (32 calls ago) has run
  double* host = R.pinned + (size_t)slot * R.pinned_doubles;
  std::memcpy(host, rq.params, (size_t)rq.n_par * sizeof(double));
  CUDA_TRY(cudaMemcpyAsync(R.d_params.p, host, (size_t)rq.n_par * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CUDA_TRY(cudaEventRecord(R.slot_done[slot], c->stream));
  CUDA_TRY(cudaGraphLaunch(R.exec, c->stream));
  c->launches += R.kernels;
  R.replays++;
  *done = true;
  return GPY_OK;
}

int evaluate_impl(gpy_context* c, const EvalRequest& rq) {
  const int D = rq.n_datasets, B = rq.n_vec;
  if (D <= 0 || B <= 0) return fail(GPY_ERR_INVALID, "need at least one dataset and one parameter vector");
  {
    bool done = false;
    GPY_TRY(evaluate_replay(c, rq, &done));
    if (done) return GPY_OK;
    c->replay.valid = false;  // whatever follows may move buffers the captured graph points into
  }
  unsigned k0_blocks = 0, k0_threads = 0;
  int k0_seg_cap = 0;
  const gpy::FluxJob* k0_jobs = nullptr;
  std::vector<int> replay_spatial_cols, replay_amp_cols;
  void (*replay_kernel)(const gpy::FoldParams) = nullptr;
  unsigned replay_grid = 0;
  if (c->h2d_pending) {
    CUDA_TRY(cudaEventSynchronize(c->h2d_done));
    c->h2d_pending = false;
  }
  std::vector<gpy::DatasetDev> dsd(D);
  std::vector<gpy::EvalDev> evals((size_t)D * B);
  std::vector<gpy::InstDev> insts;
  std::vector<int> inst_src;  // source index of every entry of insts
  std::vector<int> diffs;     // batched calls: see EvalDev::diff_begin
  std::vector<gpy::FluxJob> jobs;
  std::vector<size_t> job_flux_off;
  std::map<FluxKey, size_t> flux_seen;
  size_t flux_len = 0;
  int n_tiles = 0, max_inst = 1, max_tile_overlap = 0;
  int max_nT = 0, max_nR = 0, max_nRp = 0, max_G = 0, max_pdfc_rows = 1;
  bool tma_ok = true, all_compact = true, any_compact = false;
  const bool sequential = (B == 1);
  const uint64_t call_id = ++c->stamp;

  for (int d = 0; d < D; ++d) {
    gpy_dataset* ds = rq.datasets[d];
    if (!ds || ds->ctx != c) return fail(GPY_ERR_INVALID, "dataset %d does not belong to this context", d);
    for (auto& src : ds->sources)
      if (src.slots.size() > (size_t)kSlotsHardCap) {  // extras of an earlier large batch: drop the oldest
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        std::sort(src.slots.begin(), src.slots.end(), [](const Slot& a, const Slot& b) { return a.stamp > b.stamp; });
        for (size_t i = kSlotsPerSource; i < src.slots.size(); ++i) {
          src.slots[i].w.release();
          src.slots[i].conv.release();
        }
        src.slots.resize(kSlotsPerSource);
      }
    gpy::DatasetDev& dd = dsd[d];
    dd.nx = ds->geom.nx;
    dd.ny = ds->geom.ny;
    dd.nT = ds->nT;
    dd.nR = ds->nR;
    dd.nRp = ds->nRp;
    dd.P = (long long)ds->geom.nx * ds->geom.ny;
    dd.tile_begin = n_tiles;
    dd.n_tiles = (int)((dd.P + gpy::kFoldTPX - 1) / gpy::kFoldTPX);
    n_tiles += dd.n_tiles;
    dd.exposure = ds->exposure;
    dd.background = ds->background;
    dd.counts = rq.npred_out ? nullptr : ds->counts;
    dd.mask = ds->mask;
    dd.compact = ds->compact;
    dd.mask_row_bytes = (int)ds->mask_row_bytes;
    all_compact = all_compact && ds->compact;
    any_compact = any_compact || ds->compact;
    dd.weight = ds->weight;
    dd.ecenter = (const double*)ds->d_ecenter.p;
    dd.tmaps = ds->has_tmaps ? ds->d_tmaps.p : nullptr;
    if (!rq.npred_out && !ds->counts) return fail(GPY_ERR_INVALID, "dataset %d has no counts: stat undefined", d);

    for (int b = 0; b < B; ++b) {
      const double* pv = rq.params + (size_t)b * rq.n_par;
      gpy::EvalDev& ev = evals[(size_t)d * B + b];
      ev.inst_begin = (int)insts.size();
      ev.bkg_enabled = 0;
      ev.bkg_norm = 1.0;
      ev.bkg_tilt = 0.0;
      ev.bkg_ref = 1.0;
      if (ds->bkg.enabled) {
        GPY_TRY(check_par(ds->bkg.par_norm, rq.n_par, "background norm"));
        GPY_TRY(check_par(ds->bkg.par_tilt, rq.n_par, "background tilt"));
        GPY_TRY(check_par(ds->bkg.par_reference, rq.n_par, "background reference"));
        ev.bkg_enabled = 1;
        ev.bkg_norm = pv[ds->bkg.par_norm];
        ev.bkg_tilt = pv[ds->bkg.par_tilt];
        ev.bkg_ref = pv[ds->bkg.par_reference];
      }
      for (size_t si = 0; si < ds->sources.size(); ++si) {
        Source& src = ds->sources[si];
        const int nsp = n_spatial_params(src.desc.spatial_type);
        const int nsc = n_spectral_params(src.desc.spectral_type);
        double sp[6] = {0, 0, 0, 0, 0, 0}, sc[5] = {0, 0, 0, 0, 0};
        for (int k = 0; k < nsp; ++k) {
          GPY_TRY(check_par(src.desc.spatial_par[k], rq.n_par, "a spatial parameter"));
          sp[k] = pv[src.desc.spatial_par[k]];
          if (b == 0) replay_spatial_cols.push_back(src.desc.spatial_par[k]);
        }
        for (int k = 0; k < nsc; ++k) {
          GPY_TRY(check_par(src.desc.spectral_par[k], rq.n_par, "a spectral parameter"));
          sc[k] = pv[src.desc.spectral_par[k]];
        }
        // ---- MapEvaluator.needs_update (evaluator.py:128-144) on the persistent or a scratch state
        bool initialised = src.initialised, contributes = src.contributes;
        Rect rect = src.rect;
        double clon = src.cached_lon, clat = src.cached_lat;
        bool needs_update;
        if (!contributes) {
          needs_update = false;
        } else if (!initialised) {
          needs_update = true;
        } else {
          bool changed = !src.has_cached_spatial ||
                         std::memcmp(src.cached_spatial.data(), sp, nsp * sizeof(double)) != 0;
          if (!changed) {
            needs_update = false;
          } else {  // irf_position_changed (evaluator.py:465-481)
            double lon = sp[0] * kDegH, lat = sp[1] * kDegH;
            double sep = angular_separation_h(lon, lat, clon, clat);
            bool moved = sep > (evaluation_radius(src.desc.spatial_type, sp) + kCutoutMargin) * kDegH;
            if (moved) {
              clon = lon;
              clat = lat;
            }
            needs_update = moved;
          }
        }
        if (needs_update) {
          if (ds->irf.has_psf || ds->irf.has_edisp) {
            // the evaluator looks its IRFs up at the new position (evaluator.py:196-227) before sizing its cutout
            if (!sequential) {
              // A batched call does not advance the evaluators: its vectors use the IRFs of the last lookup as long
              // as they stay within that lookup's support (evaluation_radius + margin of ITS position -- the
              // reference's _cached_position starts at (0, 0), which would otherwise make every first spatial step a
              // new lookup); beyond it the point has to be evaluated as a single vector first.
              const double sep = angular_separation_h(sp[0] * kDegH, sp[1] * kDegH, src.lookup_lon, src.lookup_lat);
              if (src.n_irf_updates == 0 ||
                  sep > (evaluation_radius(src.desc.spatial_type, sp) + kCutoutMargin) * kDegH)
                return fail(GPY_ERR_UNSUPPORTED,
                            "a vector of a batched call moves a source beyond its IRF support (MapEvaluator.update): "
                            "evaluate that point as a single vector first");
            } else {
              double mlon, mlat;  // the IRF maps live in the dataset's frame
              frame_convert(src.desc.frame, ds->geom.frame, sp[0], sp[1], mlon, mlat);
              if (ds->irf.has_edisp) src.edisp_pix = edisp_pixel_at(ds->irf, mlon, mlat);
              if (ds->irf.has_psf) {  // (whatever apply_irf says: update() looks the kernel up, the width uses it)
                GPY_TRY(extract_psf(ds, src, mlon, mlat));
                for (auto& sl : src.slots) sl.valid = false;  // convolved with the previous kernel
                ds->version = ++c->stamp;
              } else {
                src.n_irf_updates++;
              }
              src.lookup_lon = sp[0] * kDegH;
              src.lookup_lat = sp[1] * kDegH;
            }
          }
          GPY_TRY(evaluator_update(ds, src, sp, rect, contributes));
          initialised = true;
        }
        if (sequential) {
          src.initialised = initialised;
          src.contributes = contributes;
          src.rect = rect;
          src.cached_lon = clon;
          src.cached_lat = clat;
          if (contributes) {  // parameters_spatial_changed(reset=True) inside compute_flux_spatial
            src.cached_spatial.assign(sp, sp + nsp);
            src.has_cached_spatial = true;
          }
        }
        if (!contributes) continue;
        if (rq.source_enable && !rq.source_enable[si]) continue;
        // _compute_npred: amplitude == 0 -> zeros (evaluator.py:385-389)
        const int amp_idx = src.desc.spectral_type == GPY_SPECTRAL_LP ? 0 : 1;
        if (b == 0 && !(rq.source_enable && !rq.source_enable[si])) replay_amp_cols.push_back(src.desc.spectral_par[amp_idx]);
        if (sc[amp_idx] == 0.0) continue;
        Slot* slot = nullptr;
        GPY_TRY(get_slot(ds, src, sp, rect, call_id, &slot));
        gpy::InstDev in;
        std::memset(&in, 0, sizeof(in));
        in.x0 = rect.x0;
        in.y0 = rect.y0;
        in.cx = rect.cx;
        in.cy = rect.cy;
        in.x0a = slot->x0a;
        in.pitch = slot->pitch;
        in.planes = slot->planes;
        in.group = (src.desc.apply_edisp || !ds->has_diag) ? 0 : 1;
        if (ds->irf.has_edisp) in.group = (src.desc.apply_edisp || !ds->has_diag) ? src.edisp_pix : -1;  // matrix code, remapped below
        in.plane_stride = slot->planes == 1 ? 0 : rect.cy * slot->pitch;
        in.conv = (const float*)(slot->planes == 1 ? slot->w.p : slot->conv.p);
        // Equal spectral tuples of one source across the vectors of a batch share one flux vector: K0 integrates it
        // once and the item signatures (prep_items_kernel) can tell that two vectors see the same source.
        FluxKey key;
        std::memset(&key, 0, sizeof(key));
        key.source = ((uint64_t)d << 32) | (uint64_t)si;
        std::memcpy(key.p, sc, sizeof(key.p));
        auto hit = B > 1 ? flux_seen.find(key) : flux_seen.end();
        if (hit != flux_seen.end()) {
          in.flux_off = (int)hit->second;
        } else {
          in.flux_off = (int)flux_len;
          if (B > 1) flux_seen.emplace(key, flux_len);
          gpy::FluxJob job;
          std::memset(&job, 0, sizeof(job));
          job.type = src.desc.spectral_type;
          job.n_bins = ds->nT;
          job.n_nodes = ds->n_nodes;
          job.sanitize = 1;
          for (int k = 0; k < 5; ++k) job.p[k] = sc[k];
          for (int k = 0; k < 5; ++k) job.pidx[k] = k < nsc ? src.desc.spectral_par[k] : -1;
          job.edges = (const double*)ds->d_edges.p;
          job.nodes = (const double*)ds->d_nodes.p;
          jobs.push_back(job);
          job_flux_off.push_back(flux_len);
          flux_len += ds->nT;
        }
        insts.push_back(in);
        inst_src.push_back((int)si);
      }
      ev.inst_count = (int)insts.size() - ev.inst_begin;
      max_inst = std::max(max_inst, ev.inst_count);
      // Sources that can overlap ONE tile (what a work descriptor must hold).  With few sources the count itself
      // is the bound; otherwise the largest number of cutouts that touch the rows a tile can span.
      if (ev.inst_count <= gpy::kFoldLMax + gpy::kDescOverflow) {
        max_tile_overlap = std::max(max_tile_overlap, ev.inst_count);
      } else {
        const int ny = ds->geom.ny, span = (gpy::kFoldTPX + ds->geom.nx - 1) / ds->geom.nx;  // extra rows a tile reaches
        std::vector<int> cover((size_t)ny + 2, 0);
        for (int i = ev.inst_begin; i < ev.inst_begin + ev.inst_count; ++i) {
          const int lo = std::max(insts[i].y0 - span, 0), hi = std::min(insts[i].y0 + insts[i].cy, ny);
          if (hi > lo) {
            cover[lo] += 1;
            cover[hi] -= 1;
          }
        }
        int run = 0;
        for (int y = 0; y < ny; ++y) {
          run += cover[y];
          max_tile_overlap = std::max(max_tile_overlap, run);
        }
      }
      ev.differs_all = 0;
      ev.diff_begin = (int)diffs.size();
      if (b > 0) {  // which sources differ from vector 0's: merge walk over the two lists (both in source order)
        const gpy::EvalDev& e0 = evals[(size_t)d * B];
        ev.differs_all = ev.bkg_enabled != e0.bkg_enabled || std::memcmp(&ev.bkg_norm, &e0.bkg_norm, 3 * sizeof(double)) != 0;
        int i = e0.inst_begin, j = ev.inst_begin;
        const int ie = i + e0.inst_count, je = j + ev.inst_count;
        while (!ev.differs_all && (i < ie || j < je)) {
          const int si = i < ie ? inst_src[i] : INT32_MAX, sj = j < je ? inst_src[j] : INT32_MAX;
          if (si == sj) {
            const gpy::InstDev &p = insts[i], &q = insts[j];
            if (p.conv != q.conv || p.flux_off != q.flux_off || p.group != q.group || p.x0 != q.x0 || p.y0 != q.y0 ||
                p.cx != q.cx || p.cy != q.cy) {
              diffs.push_back(i);
              diffs.push_back(j);
            }
            ++i;
            ++j;
          } else if (si < sj) {
            diffs.push_back(i++);
          } else {
            diffs.push_back(j++);
          }
        }
      }
      ev.diff_count = (int)diffs.size() - ev.diff_begin;
    }
    if (ds->irf.has_edisp) {
      // Position-dependent edisp: the instances carry the code of their matrix (nearest IRF pixel, -1 diagonal).  The
      // distinct codes of this call are the dataset's edisp groups; the tables are rebuilt when the set changes.
      const size_t i0 = (size_t)evals[(size_t)d * B].inst_begin;
      std::vector<int> codes;
      for (size_t i = i0; i < insts.size(); ++i)
        if (std::find(codes.begin(), codes.end(), insts[i].group) == codes.end()) codes.push_back(insts[i].group);
      std::sort(codes.begin(), codes.end());
      if (!codes.empty()) {
        if (codes != ds->group_codes) {
          const size_t plane = (size_t)ds->irf.geom.nx * ds->irf.geom.ny;
          // one strided view per group would need per-group strides: gather the matrices contiguously instead
          std::vector<std::vector<double>> gathered(codes.size());
          std::vector<const double*> mats(codes.size());
          for (size_t g = 0; g < codes.size(); ++g) {
            if (codes[g] < 0) {
              mats[g] = ds->h_diag.data();
            } else {
              gathered[g].resize((size_t)ds->nT * ds->nR);
              for (int t = 0; t < ds->nT; ++t)
                for (int r = 0; r < ds->nR; ++r)
                  gathered[g][(size_t)t * ds->nR + r] = ds->irf.edisp[((size_t)t * ds->nR + r) * plane + codes[g]];
              mats[g] = gathered[g].data();
            }
          }
          GPY_TRY(build_edisp_tables(ds, mats));
          ds->group_codes = codes;
          ds->version = ++c->stamp;
        }
        ds->groups_used = (int)codes.size();
        for (size_t i = i0; i < insts.size(); ++i)
          insts[i].group = (int)(std::find(codes.begin(), codes.end(), insts[i].group) - codes.begin());
      }
    }
    dd.n_groups = ds->groups_used;
    dd.edisp = (const double*)ds->d_edisp.p;
    dd.band = (const int*)ds->d_band.p;
    dd.pdfc = (const double*)ds->d_pdfc.p;
    dd.pdfc_rows = ds->pdfc_rows_cum[ds->groups_used - 1];
    max_pdfc_rows = std::max(max_pdfc_rows, dd.pdfc_rows);
    max_nT = std::max(max_nT, ds->nT);
    max_nR = std::max(max_nR, ds->nR);
    max_nRp = std::max(max_nRp, ds->nRp);
    max_G = std::max(max_G, ds->groups_used);
    const size_t P64 = (size_t)ds->geom.nx * ds->geom.ny;
    auto al16 = [](const void* p) { return ((uintptr_t)p & 15) == 0; };
    if (ds->geom.nx % 4 != 0 || P64 % 16 != 0 || !al16(ds->exposure) || !al16(ds->background) || !al16(ds->counts) ||
        !al16(ds->mask) || ds->weight != nullptr || P64 * (size_t)std::max(ds->nT, ds->nR) > (size_t)INT32_MAX)
      tma_ok = false;
  }
  gpy::FoldSmem lay = gpy::fold_tma_layout(max_nT, max_nR, max_nRp, max_G, max_pdfc_rows, all_compact);
  if (any_compact && !all_compact) tma_ok = false;  // mixed storage formats in one launch: the any-shape kernel
  // 64-pixel sub-tile kernel (three CTAs per SM): tensor maps everywhere, one edisp group, nT <= 64, nR <= 32
  bool t64 = c->use_t64 && tma_ok && max_G == 1 && max_nT <= 64 && max_nRp <= 32;
  for (int d = 0; d < D && t64; ++d) t64 = rq.datasets[d]->has_tmaps && rq.datasets[d]->has_tmaps64;
  const int t64_stages = c->t64_stages;
  if (t64) lay = gpy::fold_t64_layout(max_nT, max_nR, max_nRp, max_pdfc_rows, all_compact, t64_stages);
  if (lay.total > (size_t)c->max_smem_optin || c->force_generic ||
      max_tile_overlap > gpy::kFoldLMax + gpy::kDescOverflow)
    tma_ok = false;
  size_t max_fold_smem = 0;
  if (!tma_ok) {
    for (int d = 0; d < D; ++d)
      max_fold_smem = std::max(max_fold_smem, gpy::fold_smem_bytes(rq.datasets[d]->nT,
                                                                   rq.datasets[d]->nRp, max_inst));
    if (max_fold_smem > (size_t)c->max_smem_optin)
      return fail(GPY_ERR_UNSUPPORTED,
                  "energy axes / source list too large for the fused fold kernel (%zu bytes of shared memory)",
                  max_fold_smem);
  }

  // tile -> dataset table, cached per dataset list
  bool same = (int)c->tiles_for.size() == D;
  for (int d = 0; same && d < D; ++d)
    same = c->tiles_for[d] == rq.datasets[d] && c->tiles_ver[d] == rq.datasets[d]->version;
  if (!same) {
    std::vector<int> tile_ds((size_t)n_tiles);
    for (int d = 0; d < D; ++d) std::fill_n(tile_ds.begin() + dsd[d].tile_begin, dsd[d].n_tiles, d);
    GPY_TRY(upload_to(c->tiles, tile_ds.data(), tile_ds.size() * sizeof(int), c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));  // tile_ds is a stack vector
    c->tiles_for.assign(rq.datasets, rq.datasets + D);
    c->tiles_ver.resize(D);
    for (int d = 0; d < D; ++d) c->tiles_ver[d] = rq.datasets[d]->version;
  }

  // background factors: one K0 job per (dataset, vector), outputs [D * B][max_nRp] at the end of the flux buffer
  const size_t bkg_base = flux_len;
  for (int d = 0; d < D; ++d)
    for (int b = 0; b < B; ++b) {
      const gpy::EvalDev& ev = evals[(size_t)d * B + b];
      gpy::FluxJob job;
      std::memset(&job, 0, sizeof(job));
      job.type = gpy::kFluxJobBackground;
      job.n_bins = max_nRp;
      job.n_nodes = rq.datasets[d]->nR;
      job.sanitize = ev.bkg_enabled;
      job.p[0] = ev.bkg_norm;
      job.p[1] = ev.bkg_tilt;
      job.p[2] = ev.bkg_ref;
      for (int k = 0; k < 5; ++k) job.pidx[k] = -1;
      if (ev.bkg_enabled) {
        job.pidx[0] = rq.datasets[d]->bkg.par_norm;
        job.pidx[1] = rq.datasets[d]->bkg.par_tilt;
        job.pidx[2] = rq.datasets[d]->bkg.par_reference;
      }
      job.edges = (const double*)rq.datasets[d]->d_ecenter.p;
      jobs.push_back(job);
      job_flux_off.push_back(flux_len);
      flux_len += max_nRp;
    }
  // + max_nT: the fold kernel copies flux rows in units of the launch's largest true-energy axis
  GPY_TRY(c->flux.reserve((flux_len + (size_t)max_nT + 1) * sizeof(double)));
  GPY_TRY(c->partials.reserve((size_t)B * n_tiles * (gpy::kFoldThreads / 32) * sizeof(double)));
  GPY_TRY(c->stat.reserve((size_t)B * D * sizeof(double)));
  for (size_t i = 0; i < jobs.size(); ++i) jobs[i].out = (double*)c->flux.p + job_flux_off[i];

  Packer pk;
  size_t off_ds = pk.add(dsd.data(), dsd.size() * sizeof(gpy::DatasetDev));
  size_t off_ev = pk.add(evals.data(), evals.size() * sizeof(gpy::EvalDev));
  size_t off_in = pk.add(insts.data(), insts.size() * sizeof(gpy::InstDev));
  size_t off_jb = pk.add(jobs.data(), jobs.size() * sizeof(gpy::FluxJob));
  size_t off_df = pk.add(diffs.data(), diffs.size() * sizeof(int));
  // spatial chains queued by get_slot: job records + the first block of each job in the three grids
  const int n_sp = (int)c->pending.size();
  size_t off_sp = 0, off_first = 0, sp_conv_smem = 0;
  int sp_fill = 0, sp_over = 0, sp_conv = 0;
  if (n_sp) {
    std::vector<gpy::SpatialBatchJob> sj((size_t)n_sp);
    std::vector<int> first(3 * ((size_t)n_sp + 1));
    int* f_fill = first.data();
    int* f_over = f_fill + n_sp + 1;
    int* f_conv = f_over + n_sp + 1;
    for (int k = 0; k < n_sp; ++k) {
      const gpy_context::PendingSpatial& ps = c->pending[(size_t)k];
      sj[(size_t)k] = ps.job;
      f_fill[k] = sp_fill;
      f_over[k] = sp_over;
      f_conv[k] = sp_conv;
      const long long nf = (long long)ps.job.fill_gx * ps.job.j.ecy;
      const long long no = ps.oversample ? ((long long)ps.job.j.icx * ps.job.j.icy * 32 + 255) / 256 : 0;
      const long long nc = (long long)ps.job.conv_gx * ps.job.conv_gy * ps.job.n_planes;
      if (sp_fill + nf > INT32_MAX || sp_over + no > INT32_MAX || sp_conv + nc > INT32_MAX)
        return fail(GPY_ERR_UNSUPPORTED, "too many spatial cubes to build in one evaluation (%d)", n_sp);
      sp_fill += (int)nf;
      sp_over += (int)no;
      sp_conv += (int)nc;
      sp_conv_smem = std::max(sp_conv_smem, ps.conv_smem);
    }
    f_fill[n_sp] = sp_fill;
    f_over[n_sp] = sp_over;
    f_conv[n_sp] = sp_conv;
    off_sp = pk.add(sj.data(), sj.size() * sizeof(gpy::SpatialBatchJob));
    off_first = pk.add(first.data(), first.size() * sizeof(int));
  }
  GPY_TRY(ensure_pinned(c, pk.bytes.size()));
  std::memcpy(c->pinned, pk.bytes.data(), pk.bytes.size());
  GPY_TRY(c->staging.reserve(pk.bytes.size()));
  CUDA_TRY(cudaMemcpyAsync(c->staging.p, c->pinned, pk.bytes.size(), cudaMemcpyHostToDevice, c->stream));
  CUDA_TRY(cudaEventRecord(c->h2d_done, c->stream));
  c->h2d_pending = true;
  uint8_t* base = (uint8_t*)c->staging.p;

  if (n_sp) {
    const gpy::SpatialBatchJob* sj = (const gpy::SpatialBatchJob*)(base + off_sp);
    const int* first = (const int*)(base + off_first);
    gpy::spatial_fill_batch_kernel<<<(unsigned)sp_fill, 128, 0, c->stream>>>(sj, first, n_sp);
    c->launches++;
    if (sp_over) {
      // (jobs without an oversampled inner cutout own no block of this grid: the search skips their empty ranges)
      gpy::spatial_oversample_batch_kernel<<<(unsigned)sp_over, 256, 0, c->stream>>>(sj, first + n_sp + 1, n_sp);
      c->launches++;
    }
    if (sp_conv) {
      if (sp_conv_smem > c->conv_batch_smem_set) {
        CUDA_TRY(cudaFuncSetAttribute(gpy::psf_conv_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)sp_conv_smem));
        c->conv_batch_smem_set = sp_conv_smem;
      }
      gpy::psf_conv_batch_kernel<<<(unsigned)sp_conv, gpy::kConvThreads, sp_conv_smem, c->stream>>>(
          sj, first + 2 * ((size_t)n_sp + 1), n_sp);
      c->launches++;
    }
    CUDA_TRY(cudaGetLastError());
    c->pending.clear();
  }

  if (!jobs.empty()) {
    int seg_cap = 0;  // doubles of shared memory for the per-segment integrals (falls back to the serial loop)
    for (const auto& j : jobs)
      if (j.type != 0 && j.type != gpy::kFluxJobBackground) seg_cap = std::max(seg_cap, j.n_bins * (2 * j.n_nodes - 1));
    seg_cap = std::min(seg_cap, 5120);  // 40 KB
    gpy::spectral_flux_kernel<<<(unsigned)jobs.size(), seg_cap ? 256 : 64, (size_t)seg_cap * sizeof(double),
                                c->stream>>>((const gpy::FluxJob*)(base + off_jb), seg_cap, nullptr, nullptr);
    c->launches++;
    k0_blocks = (unsigned)jobs.size();
    k0_threads = seg_cap ? 256 : 64;
    k0_seg_cap = seg_cap;
    k0_jobs = (const gpy::FluxJob*)(base + off_jb);
  }
  gpy::FoldParams fp;
  std::memset(&fp, 0, sizeof(fp));
  fp.ds = (const gpy::DatasetDev*)(base + off_ds);
  fp.tile_ds = (const int*)c->tiles.p;
  fp.evals = (const gpy::EvalDev*)(base + off_ev);
  fp.insts = (const gpy::InstDev*)(base + off_in);
  fp.flux = (const double*)c->flux.p;
  fp.bkgfac = (const double*)c->flux.p + bkg_base;
  fp.partials = (double*)c->partials.p;
  fp.npred_out = rq.npred_out;
  fp.B = B;
  fp.n_tiles = n_tiles;
  fp.n_datasets = D;
  fp.include_background = rq.include_background;
  fp.max_inst = max_inst;
  fp.log_trunc = c->log_trunc;
  fp.smem_nT = max_nT;
  fp.smem_nR = max_nR;
  fp.smem_nRp = max_nRp;
  fp.smem_G = max_G;
  fp.smem_pdfc_rows = max_pdfc_rows;
  if (!c->zeros.p) {  // [0, 128) zeros, [128, 132) work counter, [256, 2304) fast_log table
    GPY_TRY(c->zeros.reserve(256 + 128 * 16 + 64 * 16));
    CUDA_TRY(cudaMemsetAsync(c->zeros.p, 0, 256, c->stream));
    static double tab[256];
    for (int i = 0; i < 128; ++i) {
      const double inv = 1.0 / (1.0 + (i + 0.5) / 128.0);
      tab[2 * i] = inv;
      tab[2 * i + 1] = -std::log(inv);
    }
    CUDA_TRY(cudaMemcpyAsync((uint8_t*)c->zeros.p + 256, tab, sizeof(tab), cudaMemcpyHostToDevice, c->stream));
    static double tab64[128];
    for (int i = 0; i < 64; ++i) {
      const double inv = 1.0 / (1.0 + (i + 0.5) / 64.0);
      tab64[2 * i] = inv;
      tab64[2 * i + 1] = -std::log(inv);
    }
    CUDA_TRY(cudaMemcpyAsync((uint8_t*)c->zeros.p + 256 + 2048, tab64, sizeof(tab64), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
  }
  fp.zeros = (const float*)c->zeros.p;
  fp.logtab = (const double2*)((const uint8_t*)c->zeros.p + 256);
  fp.logtab64 = (const double2*)((const uint8_t*)c->zeros.p + 256 + 2048);
  const size_t n_items = (size_t)B * n_tiles;
  const unsigned char* reduce_dup = nullptr;
  const int variant = rq.npred_out ? 1 : 0;
  if (tma_ok) {
    const gpy::DescLayout dl = gpy::desc_layout();
    fp.lay_es = (unsigned)lay.es;
    fp.lay_bs = (unsigned)lay.bs;
    fp.lay_cs = (unsigned)lay.cs;
    fp.lay_ms = (unsigned)lay.ms;
    fp.lay_vs = (unsigned)lay.vs;
    fp.lay_pdf = (unsigned)lay.pdf;
    fp.lay_desc = (unsigned)lay.desc;
    fp.lay_dyn = (unsigned)lay.dyn;
    fp.lay_bars = (unsigned)lay.bars;
    fp.dl_inst = dl.inst;
    fp.dl_over = dl.over;
    fp.dl_bytes = dl.bytes;
    gpy::PrepParams pp;
    pp.ds = fp.ds;
    pp.tile_ds = fp.tile_ds;
    pp.evals = fp.evals;
    pp.insts = fp.insts;
    pp.B = B;
    pp.n_tiles = n_tiles;
    // B > 1: evaluate each distinct (tile, vector) item once (see dedup_*_kernel)
    const bool dedup = B > 1 && !c->no_dedup;
    pp.work = nullptr;
    pp.n_work = nullptr;
    unsigned item_blocks = (unsigned)((n_items + 7) / 8);
    size_t n_work_host = n_items;
    if (dedup) {
      const size_t o_cnt = n_items * 4, o_off = o_cnt + (size_t)n_tiles * 4, o_n = o_off + (size_t)n_tiles * 4,
                   o_dup = o_n + 16;
      GPY_TRY(c->dedup.reserve(o_dup + n_items, c->stream));
      uint8_t* q = (uint8_t*)c->dedup.p;
      int* work = (int*)q;
      int* cnt = (int*)(q + o_cnt);
      int* off = (int*)(q + o_off);
      int* n_work = (int*)(q + o_n);
      unsigned char* dup = q + o_dup;
      const unsigned tb = (unsigned)((n_tiles + 7) / 8);
      gpy::dedup_flag_kernel<<<tb, 256, 0, c->stream>>>(fp.ds, fp.tile_ds, fp.evals, fp.insts,
                                                        (const int*)(base + off_df), B, n_tiles, dup, cnt);
      gpy::dedup_scan_kernel<<<1, 1024, 0, c->stream>>>(cnt, n_tiles, off, n_work);
      gpy::dedup_scatter_kernel<<<tb, 256, 0, c->stream>>>(dup, off, B, n_tiles, work);
      c->launches += 3;
      pp.work = fp.work = work;
      pp.n_work = fp.n_work = n_work;
      reduce_dup = dup;
      // The worklist length sizes the descriptor buffer and the two grids: one 4-byte read-back (the only host
      // synchronisation inside a batched call; B == 1 calls stay fully asynchronous).
      if (!c->nwork_pinned) CUDA_TRY(cudaMallocHost((void**)&c->nwork_pinned, 16));
      CUDA_TRY(cudaMemcpyAsync(c->nwork_pinned, n_work, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
      CUDA_TRY(cudaStreamSynchronize(c->stream));
      c->h2d_pending = false;
      n_work_host = (size_t)c->nwork_pinned[0];
      if (n_work_host < (size_t)n_tiles || n_work_host > n_items)
        return fail(GPY_ERR_CUDA, "batched worklist length %zu outside [%d, %zu]", n_work_host, n_tiles, n_items);
      item_blocks = (unsigned)((n_work_host + 7) / 8);
    }
    GPY_TRY(c->desc.reserve(n_work_host * (size_t)dl.bytes, c->stream));
    pp.desc = (unsigned char*)c->desc.p;
    // The descriptors hold nothing that changes on a steady-state step (no fluxes, no background factors): keep
    // them while the datasets, the source list, every cutout and every slot are what they were (B == 1 only).
    std::vector<uint8_t> key;
    if (!dedup) {
      auto put = [&key](const void* src, size_t n) {
        const uint8_t* q = (const uint8_t*)src;
        key.insert(key.end(), q, q + n);
      };
      const long long head[4] = {D, B, n_tiles, (long long)(uintptr_t)c->desc.p};
      put(head, sizeof(head));
      for (int d = 0; d < D; ++d) {
        const long long id[2] = {(long long)(uintptr_t)rq.datasets[d], (long long)rq.datasets[d]->version};
        put(id, sizeof(id));
      }
      for (const auto& ev : evals) put(&ev.inst_begin, 2 * sizeof(int));
      put(insts.data(), insts.size() * sizeof(gpy::InstDev));
    }
    if (dedup || c->no_desc_cache || key != c->desc_key) {
      gpy::prep_items_kernel<<<item_blocks, 256, 0, c->stream>>>(pp);
      c->launches++;
      c->desc_key.swap(key);  // empty after a batched call: the next B == 1 call rebuilds
    }
    fp.desc = (const unsigned char*)c->desc.p;
    // Items differ in cost (source-free tiles are light): with many items per CTA they are claimed dynamically.
    // (grid <= 2 CTAs per SM; 8 items per CTA at least.)
    const bool dynamic = B == 1 && n_work_host >= (size_t)16 * c->num_sms && !c->static_schedule;
    using FoldKernel = void (*)(const gpy::FoldParams);
    static const FoldKernel kTable[8] = {
        gpy::fold_cash_tma_kernel<false, false, false>, gpy::fold_cash_tma_kernel<false, true, false>,
        gpy::fold_cash_tma_kernel<true, false, false>,  gpy::fold_cash_tma_kernel<true, true, false>,
        gpy::fold_cash_tma_kernel<false, false, true>,  gpy::fold_cash_tma_kernel<false, true, true>,
        gpy::fold_cash_tma_kernel<true, false, true>,   gpy::fold_cash_tma_kernel<true, true, true>};
    static const FoldKernel kTable64[16] = {
        gpy::fold_cash_t64_kernel<false, false, false, 1>, gpy::fold_cash_t64_kernel<false, true, false, 1>,
        gpy::fold_cash_t64_kernel<true, false, false, 1>,  gpy::fold_cash_t64_kernel<true, true, false, 1>,
        gpy::fold_cash_t64_kernel<false, false, true, 1>,  gpy::fold_cash_t64_kernel<false, true, true, 1>,
        gpy::fold_cash_t64_kernel<true, false, true, 1>,   gpy::fold_cash_t64_kernel<true, true, true, 1>,
        gpy::fold_cash_t64_kernel<false, false, false, 2>, gpy::fold_cash_t64_kernel<false, true, false, 2>,
        gpy::fold_cash_t64_kernel<true, false, false, 2>,  gpy::fold_cash_t64_kernel<true, true, false, 2>,
        gpy::fold_cash_t64_kernel<false, false, true, 2>,  gpy::fold_cash_t64_kernel<false, true, true, 2>,
        gpy::fold_cash_t64_kernel<true, false, true, 2>,   gpy::fold_cash_t64_kernel<true, true, true, 2>};
    const int kslot = (t64 ? 8 * t64_stages : 0) + (all_compact ? 4 : 0) + variant * 2 + (dynamic ? 1 : 0);
    FoldKernel kern = t64 ? kTable64[kslot - 8] : kTable[kslot];
    if (lay.total > c->tma_smem_set[kslot]) {
      CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lay.total));
      CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                    cudaSharedmemCarveoutMaxShared));
      c->tma_smem_set[kslot] = lay.total;
    }
    int per_sm = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, gpy::kFoldThreads, lay.total));
    per_sm = std::max(per_sm, 1);
    if (c->max_ctas_per_sm > 0) per_sm = std::min(per_sm, c->max_ctas_per_sm);
    const unsigned grid = (unsigned)std::min(n_work_host, (size_t)per_sm * c->num_sms);
    if (dynamic) {
      fp.counter = (unsigned*)((uint8_t*)c->zeros.p + 128);
      CUDA_TRY(cudaMemsetAsync(fp.counter, 0, sizeof(unsigned), c->stream));
    }
    if (c->timing) {
      GPY_TRY(timing_collect(c));
      CUDA_TRY(cudaEventRecord(c->t_begin, c->stream));
    }
    kern<<<grid, gpy::kFoldThreads, lay.total, c->stream>>>(fp);
    replay_kernel = kern;
    replay_grid = grid;
    if (c->timing) {
      CUDA_TRY(cudaEventRecord(c->t_end, c->stream));
      c->timing_pending = true;
    }
  } else {
    if (max_fold_smem > c->fold_smem_set) {
      CUDA_TRY(cudaFuncSetAttribute(gpy::fold_cash_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)max_fold_smem));
      CUDA_TRY(cudaFuncSetAttribute(gpy::fold_cash_generic_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                    cudaSharedmemCarveoutMaxShared));
      c->fold_smem_set = max_fold_smem;
    }
    gpy::fold_cash_generic_kernel<<<(unsigned)n_items, gpy::kFoldThreads, max_fold_smem, c->stream>>>(fp);
  }
  c->launches++;
  if (!rq.npred_out) {
    gpy::reduce_stat_kernel<<<dim3(B, D), gpy::kReduceThreads, 0, c->stream>>>((const double*)c->partials.p, fp.ds, n_tiles, D,
                                                              B, reduce_dup, (double*)c->stat.p);
    c->launches++;
  }
  if (rq.joint_out && !rq.npred_out) {
    gpy::PeerReduceParams q = c->peer;
    q.n_datasets = D;
    gpy::peer_allreduce_kernel<<<1, 256, 0, c->stream>>>(q, (const double*)c->stat.p, B, rq.joint_out);
    c->launches++;
  }
  CUDA_TRY(cudaGetLastError());
  // Remember what this step launched: if the next call only moves spectral / background parameters it is replayed.
  gpy_context::Replay& R = c->replay;
  if (tma_ok && replay_kernel && B == 1 && !rq.npred_out && !rq.source_enable && !R.disabled && !c->timing) {
    R.datasets.assign(rq.datasets, rq.datasets + D);
    R.versions.resize(D);
    for (int d = 0; d < D; ++d) R.versions[d] = rq.datasets[d]->version;
    R.n_par = rq.n_par;
    R.include_background = rq.include_background;
    R.joint_out = rq.joint_out;
    R.spatial_cols.swap(replay_spatial_cols);
    R.spatial_vals.resize(R.spatial_cols.size());
    for (size_t i = 0; i < R.spatial_cols.size(); ++i) R.spatial_vals[i] = rq.params[R.spatial_cols[i]];
    R.amp_cols.swap(replay_amp_cols);
    R.amp_zero.resize(R.amp_cols.size());
    for (size_t i = 0; i < R.amp_cols.size(); ++i) R.amp_zero[i] = rq.params[R.amp_cols[i]] == 0.0;
    R.k0_blocks = k0_blocks;
    R.k0_threads = k0_threads;
    R.k0_seg_cap = k0_seg_cap;
    R.k0_jobs = k0_jobs;
    R.fold_kernel = replay_kernel;
    R.fold_grid = replay_grid;
    R.fold_smem = lay.total;
    R.fold_params = fp;
    R.n_tiles = n_tiles;
    R.exec_valid = false;
    if ((size_t)rq.n_par > R.pinned_doubles || !R.pinned) {
      if (R.pinned) cudaFreeHost(R.pinned);
      R.pinned = nullptr;
      R.pinned_doubles = std::max((size_t)rq.n_par, (size_t)64);
      CUDA_TRY(cudaMallocHost((void**)&R.pinned, R.pinned_doubles * kReplaySlots * sizeof(double)));
    }
    if (R.slot_done.empty()) {
      R.slot_done.resize(kReplaySlots);
      for (auto& e : R.slot_done) CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    GPY_TRY(R.d_params.reserve(R.pinned_doubles * sizeof(double)));
    R.valid = true;
  }
  return GPY_OK;
}

int evaluate(gpy_context* c, const EvalRequest& rq) {
  c->pending.clear();
  const int rc = evaluate_impl(c, rq);
  if (!c->pending.empty()) {  // failed before the queued spatial chains were launched: those slots hold nothing
    for (auto& ps : c->pending) ps.src->slots[ps.slot].valid = false;
    c->pending.clear();
  }
  return rc;
}

}  // namespace

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" {

const char* gpy_last_error(void) { return g_last_error.c_str(); }
int gpy_abi_version(void) { return GPY_ABI_VERSION; }

int gpy_init(int device, void* stream, gpy_context_t** out) {
  if (!out) return fail(GPY_ERR_INVALID, "ctx output pointer is NULL");
  *out = nullptr;
  int count = 0;
  cudaError_t err = cudaGetDeviceCount(&count);
  if (err != cudaSuccess || count == 0)
    return fail(GPY_ERR_CUDA, "no CUDA device available (%s): libgpy_b200 has no CPU fallback",
                err != cudaSuccess ? cudaGetErrorString(err) : "device count 0");
  if (device < 0 || device >= count) return fail(GPY_ERR_INVALID, "device %d outside [0, %d)", device, count);
  CUDA_TRY(cudaSetDevice(device));
  gpy_context* c = new gpy_context();
  c->device = device;
  if (stream) {
    c->stream = (cudaStream_t)stream;
  } else {
    cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
      delete c;
      return fail(GPY_ERR_CUDA, "cudaStreamCreate failed: %s", cudaGetErrorString(e));
    }
    c->own_stream = true;
  }
  cudaEventCreateWithFlags(&c->h2d_done, cudaEventDisableTiming);
  cudaDeviceGetAttribute(&c->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
  cudaDeviceGetAttribute(&c->num_sms, cudaDevAttrMultiProcessorCount, device);
  c->force_generic = std::getenv("GPY_FORCE_GENERIC_FOLD") != nullptr;
  if (const char* e = std::getenv("GPY_SPATIAL_BATCH")) c->batch_spatial = std::atoi(e) != 0;
  c->no_dedup = std::getenv("GPY_NO_DEDUP") != nullptr;
  c->static_schedule = std::getenv("GPY_STATIC_SCHEDULE") != nullptr;
  if (const char* e = std::getenv("GPY_FOLD_T64")) {
    c->use_t64 = std::atoi(e) != 0;
    c->t64_stages = std::atoi(e) == 2 ? 2 : 1;
  }
  if (const char* e = std::getenv("GPY_FOLD_CTAS_PER_SM")) c->max_ctas_per_sm = std::atoi(e);
  c->no_desc_cache = std::getenv("GPY_NO_DESC_CACHE") != nullptr;
  c->replay.disabled = std::getenv("GPY_NO_REPLAY") != nullptr;
  if (const char* m = std::getenv("GPY_FOLD_KERNEL")) c->fold_mode = std::atoi(m);
  if (c->fold_mode == 2) c->force_generic = true;
  {
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
      uint64_t keep = UINT64_MAX;  // slot buffers come and go with the batch size: keep them in the pool
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
  }
  c->log_trunc = std::log((double)(float)1e-25);  // Cython: log(<float>TRUNCATION_VALUE), .pyx:35
  // CUDA loads a kernel's code on its first use (lazy module loading): tens of milliseconds in the middle of a fit
  // the first time a source moves, a batched call arrives or an npred cube is asked for.  Touch every kernel now.
  {
    cudaFuncAttributes fa;
    const void* kernels[] = {
        (const void*)gpy::solid_angle_kernel,
        (const void*)gpy::spatial_fill_kernel,
        (const void*)gpy::spatial_oversample_kernel,
        (const void*)gpy::psf_conv_kernel,
        (const void*)gpy::spectral_flux_kernel,
        (const void*)gpy::prep_items_kernel,
        (const void*)gpy::dedup_flag_kernel,
        (const void*)gpy::dedup_scan_kernel,
        (const void*)gpy::dedup_scatter_kernel,
        (const void*)gpy::fold_cash_tma_kernel<false, false, false>,
        (const void*)gpy::fold_cash_tma_kernel<false, true, false>,
        (const void*)gpy::fold_cash_tma_kernel<true, false, false>,
        (const void*)gpy::fold_cash_tma_kernel<true, true, false>,
        (const void*)gpy::fold_cash_tma_kernel<false, false, true>,
        (const void*)gpy::fold_cash_tma_kernel<false, true, true>,
        (const void*)gpy::fold_cash_tma_kernel<true, false, true>,
        (const void*)gpy::fold_cash_tma_kernel<true, true, true>,
        (const void*)gpy::fold_cash_t64_kernel<false, false, false, 1>,
        (const void*)gpy::fold_cash_t64_kernel<false, false, false, 2>,
        (const void*)gpy::fold_cash_t64_kernel<false, true, false, 1>,
        (const void*)gpy::fold_cash_t64_kernel<false, true, false, 2>,
        (const void*)gpy::fold_cash_t64_kernel<true, false, false, 1>,
        (const void*)gpy::fold_cash_t64_kernel<true, false, false, 2>,
        (const void*)gpy::fold_cash_t64_kernel<true, true, false, 1>,
        (const void*)gpy::fold_cash_t64_kernel<true, true, false, 2>,
        (const void*)gpy::fold_cash_t64_kernel<false, false, true, 1>,
        (const void*)gpy::fold_cash_t64_kernel<false, false, true, 2>,
        (const void*)gpy::fold_cash_t64_kernel<false, true, true, 1>,
        (const void*)gpy::fold_cash_t64_kernel<false, true, true, 2>,
        (const void*)gpy::fold_cash_t64_kernel<true, false, true, 1>,
        (const void*)gpy::fold_cash_t64_kernel<true, false, true, 2>,
        (const void*)gpy::fold_cash_t64_kernel<true, true, true, 1>,
        (const void*)gpy::fold_cash_t64_kernel<true, true, true, 2>,
        (const void*)gpy::fold_cash_generic_kernel,
        (const void*)gpy::reduce_stat_kernel,
        (const void*)gpy::peer_allreduce_kernel,
        (const void*)gpy::cash_sum_kernel,
        (const void*)gpy::wstat_sum_kernel<float>,
        (const void*)gpy::wstat_sum_kernel<double>,
        (const void*)gpy::final_sum_kernel,
        (const void*)gpy::ts_map_kernel,
        (const void*)gpy::ts_map_strip_kernel,
        (const void*)gpy::f_cash_root_kernel,
        (const void*)gpy::norm_bounds_kernel,
    };
    for (const void* k : kernels) (void)cudaFuncGetAttributes(&fa, k);
    (void)cudaGetLastError();
  }
  *out = c;
  return GPY_OK;
}

int gpy_finalize(gpy_context_t* c) {
  if (!c) return GPY_OK;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  for (DevBuf* b : {&c->peer_seq, &c->zeros, &c->desc, &c->dedup, &c->staging, &c->flux, &c->partials, &c->stat, &c->tiles, &c->scratch_a, &c->scratch_b, &c->scratch_c})
    b->release();
  if (c->pinned) cudaFreeHost(c->pinned);
  if (c->replay.exec) cudaGraphExecDestroy(c->replay.exec);
  if (c->replay.pinned) cudaFreeHost(c->replay.pinned);
  for (cudaEvent_t e : c->replay.slot_done) cudaEventDestroy(e);
  c->replay.d_params.release();
  if (c->stat_pinned) cudaFreeHost(c->stat_pinned);
  if (c->nwork_pinned) cudaFreeHost(c->nwork_pinned);
  if (c->h2d_done) cudaEventDestroy(c->h2d_done);
  if (c->t_begin) cudaEventDestroy(c->t_begin);
  if (c->t_end) cudaEventDestroy(c->t_end);
  if (c->own_stream) cudaStreamDestroy(c->stream);
  delete c;
  return GPY_OK;
}

int64_t gpy_launch_count(const gpy_context_t* c) { return c ? c->launches : 0; }

int gpy_kernel_timing(gpy_context_t* c, int32_t enable) {
  if (!c) return fail(GPY_ERR_INVALID, "NULL context");
  CUDA_TRY(cudaSetDevice(c->device));
  if (enable && !c->t_begin) {
    CUDA_TRY(cudaEventCreate(&c->t_begin));
    CUDA_TRY(cudaEventCreate(&c->t_end));
  }
  if (enable) {
    c->timing_ms = 0.0;
    c->timing_launches = 0;
    c->timing_pending = false;
  }
  c->timing = enable != 0;
  return GPY_OK;
}

int gpy_kernel_time(gpy_context_t* c, double* total_ms, int64_t* launches) {
  if (!c || !total_ms || !launches) return fail(GPY_ERR_INVALID, "NULL argument to gpy_kernel_time");
  CUDA_TRY(cudaSetDevice(c->device));
  GPY_TRY(timing_collect(c));
  *total_ms = c->timing_ms;
  *launches = c->timing_launches;
  return GPY_OK;
}

int gpy_dataset_create(gpy_context_t* c, const gpy_dataset_desc* desc, gpy_dataset_t** out) {
  if (!c || !desc || !out) return fail(GPY_ERR_INVALID, "NULL argument to gpy_dataset_create");
  *out = nullptr;
  const gpy_geom_desc& g = desc->geom;
  if (g.proj != 0 && g.proj != 1) return fail(GPY_ERR_UNSUPPORTED, "projection code %d (0 CAR, 1 TAN)", g.proj);
  if (!valid_frame(g.frame)) return fail(GPY_ERR_UNSUPPORTED, "celestial frame code %d (0 galactic, 1 icrs)", g.frame);
  if (g.proj == 0 && !(g.crval_lat > -90.0 && g.crval_lat < 90.0))
    return fail(GPY_ERR_INVALID, "CAR: reference latitude %g outside (-90, 90) deg", g.crval_lat);
  if (g.nx <= 0 || g.ny <= 0 || g.n_true <= 0 || g.n_reco <= 0 || !(g.binsz > 0))
    return fail(GPY_ERR_INVALID, "invalid geometry (nx=%d ny=%d n_true=%d n_reco=%d binsz=%g)", g.nx, g.ny, g.n_true,
                g.n_reco, g.binsz);
  if (!desc->exposure || !desc->edisp || !desc->energy_true_edges || !desc->energy_reco_center)
    return fail(GPY_ERR_INVALID, "exposure, edisp and the energy axes are required");
  if (((uintptr_t)desc->exposure & 15) != 0)
    return fail(GPY_ERR_INVALID, "exposure must be 16-byte aligned");
  if (desc->psf_kernel && (desc->psf_k <= 0 || desc->psf_k % 2 == 0))
    return fail(GPY_ERR_INVALID, "psf_k must be odd and positive, got %d", desc->psf_k);
  if (desc->n_nodes < 2 || !desc->nodes) return fail(GPY_ERR_INVALID, "integrate_spectrum nodes missing (n_nodes >= 2)");
  CUDA_TRY(cudaSetDevice(c->device));
  gpy_dataset* ds = new gpy_dataset();
  ds->ctx = c;
  ds->geom = geom_from_desc(g);
  ds->nT = g.n_true;
  ds->nR = g.n_reco;
  ds->nRp = (g.n_reco + 7) / 8 * 8;
  ds->n_nodes = desc->n_nodes;
  ds->exposure = desc->exposure;
  ds->background = desc->background;
  ds->counts = desc->counts;
  ds->mask = desc->mask;
  ds->compact = desc->compact;
  ds->mask_row_bytes = desc->mask_row_bytes;
  if (ds->compact) {
    const bool bad_rows = desc->mask && (desc->mask_row_bytes % 16 != 0 || desc->mask_row_bytes * 8 < (int64_t)g.nx * g.ny);
    if (ds->compact != 1 || desc->weight || bad_rows) {
      delete ds;
      if (desc->weight) return fail(GPY_ERR_UNSUPPORTED, "float weights (cash_weighted) with the compact storage format");
      return fail(GPY_ERR_INVALID, "compact must be 0 or 1; compact mask rows must be a multiple of 16 bytes and hold "
                  "ny * nx bits (compact %d, mask_row_bytes %lld)", desc->compact, (long long)desc->mask_row_bytes);
    }
  }
  ds->weight = desc->weight;
  ds->version = ++c->stamp;
  const size_t P = (size_t)g.nx * g.ny;
  if (desc->mask_image) ds->mask_image.assign(desc->mask_image, desc->mask_image + P);
  ds->has_diag = desc->edisp_diagonal != nullptr;
  ds->n_groups = ds->has_diag ? 2 : 1;
  auto cleanup = [&](int rc) {
    gpy_dataset_destroy(ds);
    return rc;
  };
  cudaStream_t s = c->stream;
  int rc;
  if ((rc = upload_to(ds->d_edges, desc->energy_true_edges, (ds->nT + 1) * sizeof(double), s))) return cleanup(rc);
  if ((rc = upload_to(ds->d_ecenter, desc->energy_reco_center, ds->nR * sizeof(double), s))) return cleanup(rc);
  if ((rc = upload_to(ds->d_nodes, desc->nodes, (size_t)ds->nT * ds->n_nodes * sizeof(double), s))) return cleanup(rc);
  ds->h_edisp.assign(desc->edisp, desc->edisp + (size_t)ds->nT * ds->nR);
  ds->group_codes.assign(1, -2);
  std::vector<const double*> mats{ds->h_edisp.data()};
  if (ds->has_diag) {
    ds->h_diag.assign(desc->edisp_diagonal, desc->edisp_diagonal + (size_t)ds->nT * ds->nR);
    ds->group_codes.push_back(-1);
    mats.push_back(ds->h_diag.data());
  }
  if ((rc = build_edisp_tables(ds, mats))) return cleanup(rc);
  std::vector<float> kflip;
  if (desc->psf_kernel) {
    ds->has_psf = true;
    ds->psf_k = desc->psf_k;
    prepare_planes(desc->psf_kernel, ds->nT, desc->psf_k, kflip, ds->planes, ds->max_h, ds->max_kw);
    if ((rc = upload_to(ds->d_kflip, kflip.data(), kflip.size() * sizeof(float), s))) return cleanup(rc);
    if ((rc = upload_to(ds->d_planes, ds->planes.data(), ds->planes.size() * sizeof(gpy::ConvPlane), s)))
      return cleanup(rc);
  }
  if ((rc = ds->d_omega.reserve(P * sizeof(double)))) return cleanup(rc);
  gpy::solid_angle_kernel<<<dim3((g.nx + 127) / 128, g.ny), 128, 0, s>>>(wcs_of(ds->geom), g.nx, g.ny,
                                                                          (double*)ds->d_omega.p);
  c->launches++;
  cudaError_t e = cudaStreamSynchronize(s);  // host staging vectors go out of scope
  if (e != cudaSuccess) return cleanup(fail(GPY_ERR_CUDA, "dataset upload failed: %s", cudaGetErrorString(e)));
  if ((rc = build_tmaps(ds))) return cleanup(rc);
  *out = ds;
  return GPY_OK;
}

int gpy_dataset_destroy(gpy_dataset_t* ds) {
  if (!ds) return GPY_OK;
  cudaSetDevice(ds->ctx->device);
  cudaStreamSynchronize(ds->ctx->stream);
  for (DevBuf* b : {&ds->d_edges, &ds->d_nodes, &ds->d_ecenter, &ds->d_edisp, &ds->d_band, &ds->d_pdfc, &ds->d_tmaps, &ds->d_omega, &ds->d_kflip,
                    &ds->d_planes})
    b->release();
  ds->irf.d_rad_center.release();
  for (auto& src : ds->sources) src.release_irf();
  for (auto& src : ds->sources)
    for (auto& sl : src.slots) {
      sl.w.release();
      sl.conv.release();
    }
  for (auto& b : ds->retired_tmaps) b.release();
  ds->ctx->tiles_for.clear();
  delete ds;
  return GPY_OK;
}

int gpy_dataset_set_counts(gpy_dataset_t* ds, const float* counts_device) {
  if (!ds) return fail(GPY_ERR_INVALID, "NULL dataset");
  ds->counts = counts_device;
  cudaStreamSynchronize(ds->ctx->stream);
  return build_tmaps(ds);
}

int gpy_dataset_set_models(gpy_dataset_t* ds, const gpy_source_desc* sources, int32_t n_sources,
                           const gpy_background_desc* background) {
  if (!ds) return fail(GPY_ERR_INVALID, "NULL dataset");
  if (n_sources < 0 || (n_sources > 0 && !sources)) return fail(GPY_ERR_INVALID, "bad source list");
  for (int i = 0; i < n_sources; ++i) {
    if (sources[i].spatial_type < 0 || sources[i].spatial_type > 2)
      return fail(GPY_ERR_UNSUPPORTED, "source %d: unsupported spatial model type %d", i, sources[i].spatial_type);
    if (sources[i].spectral_type < 0 || sources[i].spectral_type > 2)
      return fail(GPY_ERR_UNSUPPORTED, "source %d: unsupported spectral model type %d", i, sources[i].spectral_type);
    if (!valid_frame(sources[i].frame))
      return fail(GPY_ERR_UNSUPPORTED, "source %d: celestial frame code %d (0 galactic, 1 icrs)", i, sources[i].frame);
  }
  cudaSetDevice(ds->ctx->device);
  cudaStreamSynchronize(ds->ctx->stream);
  for (auto& src : ds->sources) {
    for (auto& sl : src.slots) {
      sl.w.release();
      sl.conv.release();
    }
    src.release_irf();
  }
  ds->sources.clear();
  ds->sources.resize(n_sources);
  ds->groups_used = 1;
  for (int i = 0; i < n_sources; ++i) {
    ds->sources[i].desc = sources[i];
    if (!sources[i].apply_edisp && ds->has_diag) ds->groups_used = 2;
  }
  if (background)
    ds->bkg = *background;
  else
    ds->bkg = gpy_background_desc{0, -1, -1, -1};
  return GPY_OK;
}

int gpy_dataset_set_irf_maps(gpy_dataset_t* ds, const gpy_irf_maps_desc* desc) {
  if (!ds || !desc) return fail(GPY_ERR_INVALID, "NULL argument to gpy_dataset_set_irf_maps");
  if (!desc->psf_map && !desc->edisp_map) return fail(GPY_ERR_INVALID, "neither a PSF map nor an edisp map given");
  if (desc->nx <= 0 || desc->ny <= 0 || !(desc->binsz > 0))
    return fail(GPY_ERR_INVALID, "bad IRF grid %d x %d, binsz %g", desc->nx, desc->ny, desc->binsz);
  if (desc->proj != 0 && desc->proj != 1) return fail(GPY_ERR_UNSUPPORTED, "projection code %d (0 CAR, 1 TAN)", desc->proj);
  if (desc->psf_map && (desc->n_rad < 2 || !desc->rad_edges))
    return fail(GPY_ERR_INVALID, "a PSF map needs rad_edges with at least two bins");
  gpy_context* c = ds->ctx;
  CUDA_TRY(cudaSetDevice(c->device));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  IrfMaps& m = ds->irf;
  m.geom = GeomH{desc->nx, desc->ny, desc->binsz, desc->crval_lon, desc->crpix_x, desc->crpix_y};
  m.geom.crval_lat = desc->crval_lat;
  m.geom.proj = desc->proj;
  const size_t plane = (size_t)desc->nx * desc->ny;
  m.has_psf = desc->psf_map != nullptr;
  m.has_edisp = desc->edisp_map != nullptr;
  if (m.has_psf) {
    m.n_rad = desc->n_rad;
    m.rad_edges.assign(desc->rad_edges, desc->rad_edges + desc->n_rad + 1);
    m.rad_center.resize(desc->n_rad);
    for (int k = 0; k < desc->n_rad; ++k) {
      if (!(m.rad_edges[k + 1] > m.rad_edges[k])) return fail(GPY_ERR_INVALID, "rad_edges must increase");
      m.rad_center[k] = 0.5 * (m.rad_edges[k] + m.rad_edges[k + 1]);
    }
    m.psf.assign(desc->psf_map, desc->psf_map + (size_t)ds->nT * desc->n_rad * plane);
    GPY_TRY(upload_to(m.d_rad_center, m.rad_center.data(), m.rad_center.size() * sizeof(double), c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    ds->has_psf = true;
  }
  if (m.has_edisp) m.edisp.assign(desc->edisp_map, desc->edisp_map + (size_t)ds->nT * ds->nR * plane);
  // every evaluator looks its IRFs up again at its next evaluation (as after MapDataset.models = ...)
  for (auto& src : ds->sources) {
    src.initialised = false;
    src.contributes = true;
    src.has_cached_spatial = false;
    src.cached_lon = src.cached_lat = 0.0;
    for (auto& sl : src.slots) sl.valid = false;
  }
  ds->version = ++c->stamp;
  return GPY_OK;
}

int gpy_dataset_source_psf_kernel(const gpy_dataset_t* ds, int32_t source, float* kernel, int32_t capacity, int32_t* k,
                                  int32_t* edisp_pixel, int32_t* n_updates) {
  if (!ds || !k) return fail(GPY_ERR_INVALID, "NULL argument");
  if (source < 0 || source >= (int)ds->sources.size()) return fail(GPY_ERR_INVALID, "source index out of range");
  const Source& s = ds->sources[source];
  *k = ds->irf.has_psf ? s.psf_k : ds->psf_k;
  if (edisp_pixel) *edisp_pixel = ds->irf.has_edisp ? s.edisp_pix : -1;
  if (n_updates) *n_updates = s.n_irf_updates;
  if (kernel) {
    if (!ds->irf.has_psf || !s.p_kernel.p)
      return fail(GPY_ERR_INVALID, "source %d has no position-dependent PSF kernel (yet)", source);
    const size_t n = (size_t)ds->nT * s.psf_k * s.psf_k;
    if ((size_t)capacity < n) return fail(GPY_ERR_INVALID, "kernel buffer too small: %d < %zu floats", capacity, n);
    CUDA_TRY(cudaSetDevice(ds->ctx->device));
    CUDA_TRY(cudaMemcpyAsync(kernel, s.p_kernel.p, n * sizeof(float), cudaMemcpyDeviceToHost, ds->ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ds->ctx->stream));
  }
  return GPY_OK;
}

int gpy_dataset_source_state(const gpy_dataset_t* ds, int32_t source, gpy_source_state* out) {
  if (!ds || !out) return fail(GPY_ERR_INVALID, "NULL argument");
  if (source < 0 || source >= (int)ds->sources.size()) return fail(GPY_ERR_INVALID, "source index out of range");
  const Source& s = ds->sources[source];
  out->initialised = s.initialised;
  out->contributes = s.contributes;
  out->x0 = s.rect.x0;
  out->y0 = s.rect.y0;
  out->cx = s.rect.cx;
  out->cy = s.rect.cy;
  out->oversampling = s.last_f;
  out->n_spatial_evals = s.n_spatial_evals;
  return GPY_OK;
}

int gpy_stat_sum_device(gpy_context_t* c, gpy_dataset_t* const* datasets, int32_t n_datasets, const double* params,
                        int32_t n_vec, int32_t n_par, double* stat_device) {
  if (!c || !datasets || !params || !stat_device) return fail(GPY_ERR_INVALID, "NULL argument to gpy_stat_sum_device");
  CUDA_TRY(cudaSetDevice(c->device));
  EvalRequest rq{datasets, n_datasets, params, n_vec, n_par, nullptr, 1, nullptr};
  GPY_TRY(evaluate(c, rq));
  CUDA_TRY(cudaMemcpyAsync(stat_device, c->stat.p, (size_t)n_vec * n_datasets * sizeof(double),
                           cudaMemcpyDeviceToDevice, c->stream));
  return GPY_OK;
}

int gpy_peer_reduce_init(gpy_context_t* c, void* const* peer_buffers, int32_t rank, int32_t world, int32_t capacity) {
  if (!c || !peer_buffers) return fail(GPY_ERR_INVALID, "NULL argument to gpy_peer_reduce_init");
  if (world < 1 || world > gpy::kPeerMaxWorld || rank < 0 || rank >= world || capacity < 1)
    return fail(GPY_ERR_INVALID, "bad rank / world / capacity (%d / %d / %d)", rank, world, capacity);
  CUDA_TRY(cudaSetDevice(c->device));
  for (int r = 0; r < world; ++r) {
    if (!peer_buffers[r]) return fail(GPY_ERR_INVALID, "peer buffer %d is NULL", r);
    c->peer.bufs[r] = (double*)peer_buffers[r];
  }
  c->peer.rank = rank;
  c->peer.world = world;
  c->peer.cap = capacity;
  GPY_TRY(c->peer_seq.reserve(sizeof(unsigned long long)));
  CUDA_TRY(cudaMemsetAsync(c->peer_seq.p, 0, sizeof(unsigned long long), c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  c->peer.seq = (unsigned long long*)c->peer_seq.p;
  c->peer_on = true;
  return GPY_OK;
}

int64_t gpy_peer_reduce_bytes(int32_t world, int32_t capacity) {
  return (int64_t)2 * world * ((int64_t)capacity + 1) * (int64_t)sizeof(double);
}

int gpy_stat_sum_joint_device(gpy_context_t* c, gpy_dataset_t* const* datasets, int32_t n_datasets, const double* params,
                              int32_t n_vec, int32_t n_par, double* joint_device) {
  if (!c || !datasets || !params || !joint_device) return fail(GPY_ERR_INVALID, "NULL argument to gpy_stat_sum_joint_device");
  if (!c->peer_on) return fail(GPY_ERR_INVALID, "gpy_peer_reduce_init has not been called on this context");
  if (n_vec > c->peer.cap) return fail(GPY_ERR_INVALID, "%d parameter vectors exceed the peer-reduce capacity %d", n_vec, c->peer.cap);
  CUDA_TRY(cudaSetDevice(c->device));
  EvalRequest rq{datasets, n_datasets, params, n_vec, n_par, nullptr, 1, nullptr};
  rq.joint_out = joint_device;  // the all-reduce is the last kernel of the evaluation (inside the replayed graph too)
  GPY_TRY(evaluate(c, rq));
  return GPY_OK;
}

int gpy_stat_sum_joint(gpy_context_t* c, gpy_dataset_t* const* datasets, int32_t n_datasets, const double* params,
                       int32_t n_vec, int32_t n_par, double* joint) {
  if (!c || !joint) return fail(GPY_ERR_INVALID, "NULL argument to gpy_stat_sum_joint");
  if (!c->peer_on) return fail(GPY_ERR_INVALID, "gpy_peer_reduce_init has not been called on this context");
  CUDA_TRY(cudaSetDevice(c->device));
  // one device buffer for the life of the context: its address is part of the replayed graph
  if (!c->joint_dev.p) GPY_TRY(c->joint_dev.reserve((size_t)std::max(c->peer.cap, 1) * sizeof(double)));
  GPY_TRY(gpy_stat_sum_joint_device(c, datasets, n_datasets, params, n_vec, n_par, (double*)c->joint_dev.p));
  GPY_TRY(ensure_stat_pinned(c, (size_t)n_vec));
  CUDA_TRY(cudaMemcpyAsync(c->stat_pinned, c->joint_dev.p, (size_t)n_vec * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  c->h2d_pending = false;
  std::memcpy(joint, c->stat_pinned, (size_t)n_vec * sizeof(double));
  return GPY_OK;
}

int gpy_stat_sum(gpy_context_t* c, gpy_dataset_t* const* datasets, int32_t n_datasets, const double* params,
                 int32_t n_vec, int32_t n_par, double* stat, double* stat_per_dataset) {
  if (!c || !datasets || !params || !stat) return fail(GPY_ERR_INVALID, "NULL argument to gpy_stat_sum");
  CUDA_TRY(cudaSetDevice(c->device));
  EvalRequest rq{datasets, n_datasets, params, n_vec, n_par, nullptr, 1, nullptr};
  GPY_TRY(evaluate(c, rq));
  const size_t n = (size_t)n_vec * n_datasets;
  GPY_TRY(ensure_stat_pinned(c, n));
  CUDA_TRY(cudaMemcpyAsync(c->stat_pinned, c->stat.p, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  c->h2d_pending = false;
  for (int b = 0; b < n_vec; ++b) {
    double s = 0.0;  // Datasets.stat_sum: python loop `stat_sum += dataset.stat_sum_likelihood()` in dataset order
    for (int d = 0; d < n_datasets; ++d) s += c->stat_pinned[(size_t)b * n_datasets + d];
    stat[b] = s;
  }
  if (stat_per_dataset) std::memcpy(stat_per_dataset, c->stat_pinned, n * sizeof(double));
  return GPY_OK;
}

int gpy_npred(gpy_context_t* c, gpy_dataset_t* ds, const double* params, int32_t n_par, const uint8_t* source_enable,
              int32_t include_background, double* npred_device) {
  if (!c || !ds || !params || !npred_device) return fail(GPY_ERR_INVALID, "NULL argument to gpy_npred");
  CUDA_TRY(cudaSetDevice(c->device));
  gpy_dataset_t* list[1] = {ds};
  EvalRequest rq{list, 1, params, 1, n_par, source_enable, include_background, npred_device};
  GPY_TRY(evaluate(c, rq));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  c->h2d_pending = false;
  return GPY_OK;
}

static int cash_common(gpy_context_t* c, const double* counts, const double* npred, const double* weight, int64_t n,
                       double* out) {
  if (!c || !out || (n > 0 && (!counts || !npred))) return fail(GPY_ERR_INVALID, "NULL argument to gpy_cash_sum");
  if (n < 0) return fail(GPY_ERR_INVALID, "negative length");
  if (n == 0) {
    *out = 0.0;
    return GPY_OK;
  }
  CUDA_TRY(cudaSetDevice(c->device));
  const size_t bytes = (size_t)n * sizeof(double);
  GPY_TRY(c->scratch_a.reserve(bytes));
  GPY_TRY(c->scratch_b.reserve(bytes));
  if (weight) GPY_TRY(c->scratch_c.reserve(bytes));
  CUDA_TRY(cudaMemcpyAsync(c->scratch_a.p, counts, bytes, cudaMemcpyHostToDevice, c->stream));
  CUDA_TRY(cudaMemcpyAsync(c->scratch_b.p, npred, bytes, cudaMemcpyHostToDevice, c->stream));
  if (weight) CUDA_TRY(cudaMemcpyAsync(c->scratch_c.p, weight, bytes, cudaMemcpyHostToDevice, c->stream));
  int blocks = (int)std::min<int64_t>((n + 255) / 256, 148 * 8);
  GPY_TRY(c->partials.reserve((size_t)blocks * sizeof(double)));
  GPY_TRY(c->stat.reserve(sizeof(double)));
  gpy::cash_sum_kernel<<<blocks, 256, 0, c->stream>>>((const double*)c->scratch_a.p, (const double*)c->scratch_b.p,
                                                       weight ? (const double*)c->scratch_c.p : nullptr, n, c->log_trunc,
                                                       (double*)c->partials.p);
  gpy::final_sum_kernel<<<1, 256, 0, c->stream>>>((const double*)c->partials.p, blocks, 2.0, (double*)c->stat.p);
  c->launches += 2;
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(out, c->stat.p, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  return GPY_OK;
}

int gpy_cash_sum(gpy_context_t* c, const double* counts, const double* npred, int64_t n, double* out) {
  return cash_common(c, counts, npred, nullptr, n, out);
}

int gpy_weighted_cash_sum(gpy_context_t* c, const double* counts, const double* npred, const double* weight, int64_t n,
                          double* out) {
  if (!weight && n > 0) return fail(GPY_ERR_INVALID, "NULL weight");
  return cash_common(c, counts, npred, weight, n, out);
}

// WStatFitStatistic.stat_sum_dataset (stats/fit_statistics.py:332-361) on DEVICE arrays.
int gpy_wstat_sum_device(gpy_context_t* c, const void* n_on, const void* n_off, const void* alpha, int32_t float64_inputs,
                         const double* mu_sig, const uint8_t* mask, int64_t n, double* stat_per_bin, double* out) {
  if (!c || !out) return fail(GPY_ERR_INVALID, "NULL argument to gpy_wstat_sum_device");
  if (n < 0) return fail(GPY_ERR_INVALID, "negative length");
  if (n > 0 && (!n_on || !n_off || !alpha || !mu_sig)) return fail(GPY_ERR_INVALID, "NULL array");
  CUDA_TRY(cudaSetDevice(c->device));
  const int blocks = (int)std::min<int64_t>(std::max<int64_t>((n + 255) / 256, 1), (int64_t)c->num_sms * 8);
  GPY_TRY(c->partials.reserve((size_t)blocks * sizeof(double), c->stream));
  GPY_TRY(c->stat.reserve(4 * sizeof(double)));
  if (float64_inputs)
    gpy::wstat_sum_kernel<double><<<blocks, 256, 0, c->stream>>>((const double*)n_on, (const double*)n_off,
                                                                (const double*)alpha, mu_sig, mask, n, stat_per_bin,
                                                                (double*)c->partials.p);
  else
    gpy::wstat_sum_kernel<float><<<blocks, 256, 0, c->stream>>>((const float*)n_on, (const float*)n_off,
                                                               (const float*)alpha, mu_sig, mask, n, stat_per_bin,
                                                               (double*)c->partials.p);
  gpy::final_sum_kernel<<<1, 256, 0, c->stream>>>((const double*)c->partials.p, blocks, 1.0, (double*)c->stat.p);
  c->launches += 2;
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(out, c->stat.p, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  return GPY_OK;
}

static int registry3(gpy_context_t* c, const double* counts, const double* background, const double* model,
                     int64_t n, double** d_c, double** d_b, double** d_m) {
  if (n < 0) return fail(GPY_ERR_INVALID, "negative length");
  if (n > 0 && (!counts || !background || !model)) return fail(GPY_ERR_INVALID, "NULL array");
  CUDA_TRY(cudaSetDevice(c->device));
  const size_t bytes = (size_t)std::max<int64_t>(n, 1) * sizeof(double);
  GPY_TRY(c->scratch_a.reserve(bytes));
  GPY_TRY(c->scratch_b.reserve(bytes));
  GPY_TRY(c->scratch_c.reserve(bytes));
  GPY_TRY(c->stat.reserve(4 * sizeof(double)));
  if (n > 0) {
    CUDA_TRY(cudaMemcpyAsync(c->scratch_a.p, counts, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(c->scratch_b.p, background, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(c->scratch_c.p, model, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
  }
  *d_c = (double*)c->scratch_a.p;
  *d_b = (double*)c->scratch_b.p;
  *d_m = (double*)c->scratch_c.p;
  return GPY_OK;
}

int gpy_f_cash_root(gpy_context_t* c, double x, const double* counts, const double* background, const double* model,
                    int64_t n, double* out) {
  if (!c || !out) return fail(GPY_ERR_INVALID, "NULL argument to gpy_f_cash_root");
  double *d_c, *d_b, *d_m;
  GPY_TRY(registry3(c, counts, background, model, n, &d_c, &d_b, &d_m));
  gpy::f_cash_root_kernel<<<1, 256, 0, c->stream>>>(x, d_c, d_b, d_m, n, (double*)c->stat.p);
  c->launches++;
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(out, c->stat.p, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  return GPY_OK;
}

int gpy_norm_bounds(gpy_context_t* c, const double* counts, const double* background, const double* model, int64_t n,
                    double* out3) {
  if (!c || !out3) return fail(GPY_ERR_INVALID, "NULL argument to gpy_norm_bounds");
  double *d_c, *d_b, *d_m;
  GPY_TRY(registry3(c, counts, background, model, n, &d_c, &d_b, &d_m));
  gpy::norm_bounds_kernel<<<1, 256, 0, c->stream>>>(d_c, d_b, d_m, n, (double*)c->stat.p);
  c->launches++;
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(out3, c->stat.p, 3 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  return GPY_OK;
}

int gpy_ts_map(gpy_context_t* c, const double* counts, const double* background, const double* exposure,
               int32_t n_planes, int32_t ny, int32_t nx, const double* kernel, int32_t ky, int32_t kx,
               const uint8_t* mask, double rtol, int32_t max_niter, double n_sigma, double* out) {
  if (!c || !counts || !background || !exposure || !kernel || !out)
    return fail(GPY_ERR_INVALID, "NULL argument to gpy_ts_map");
  if (n_planes <= 0 || ny <= 0 || nx <= 0 || ky <= 0 || kx <= 0 || ky % 2 == 0 || kx % 2 == 0)
    return fail(GPY_ERR_INVALID, "bad shape for gpy_ts_map (kernel sides must be odd)");
  if (max_niter < 0 || !(rtol >= 4 * std::numeric_limits<double>::epsilon()))
    return fail(GPY_ERR_INVALID, "rtol too small (%g) or negative maxiter", rtol);  // scipy: ValueError
  CUDA_TRY(cudaSetDevice(c->device));
  const size_t P = (size_t)ny * nx, cube = P * n_planes * sizeof(double);
  const size_t kbytes = (size_t)n_planes * ky * kx * sizeof(double);
  DevBuf d_in, d_k, d_mask, d_out;
  int rc = GPY_OK;
  auto done = [&](int r) {
    d_in.release();
    d_k.release();
    d_mask.release();
    d_out.release();
    return r;
  };
  if ((rc = d_in.reserve(3 * cube))) return done(rc);
  if ((rc = d_k.reserve(kbytes))) return done(rc);
  if ((rc = d_out.reserve(GPY_TS_PLANES * P * sizeof(double)))) return done(rc);
  if (mask && (rc = d_mask.reserve(P))) return done(rc);
  uint8_t* q = (uint8_t*)d_in.p;
  cudaError_t e = cudaMemcpyAsync(q, counts, cube, cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(q + cube, background, cube, cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(q + 2 * cube, exposure, cube, cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_k.p, kernel, kbytes, cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess && mask) e = cudaMemcpyAsync(d_mask.p, mask, P, cudaMemcpyHostToDevice, c->stream);
  if (e != cudaSuccess) return done(fail(GPY_ERR_CUDA, "TS-map upload failed: %s", cudaGetErrorString(e)));
  gpy::TsMapParams tp;
  tp.counts = (const double*)q;
  tp.background = (const double*)(q + cube);
  tp.exposure = (const double*)(q + 2 * cube);
  tp.kernel = (const double*)d_k.p;
  tp.mask = mask ? (const unsigned char*)d_mask.p : nullptr;
  tp.out = (double*)d_out.p;
  tp.n_planes = n_planes;
  tp.ny = ny;
  tp.nx = nx;
  tp.ky = ky;
  tp.kx = kx;
  tp.max_niter = max_niter;
  tp.rtol = rtol;
  tp.xtol = 2e-12;
  tp.n_sigma = n_sigma;
  tp.log_trunc = c->log_trunc;
  const size_t strip_smem = gpy::ts_strip_smem_bytes(n_planes, ky, kx);
  if (strip_smem <= 100 * 1024 && !std::getenv("GPY_TS_GLOBAL")) {  // >= 2 blocks per SM; else read through L1 / L2
    if (strip_smem > c->ts_smem_set) {
      cudaError_t ea = cudaFuncSetAttribute(gpy::ts_map_strip_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)strip_smem);
      if (ea != cudaSuccess) return done(fail(GPY_ERR_CUDA, "TS-map shared memory: %s", cudaGetErrorString(ea)));
      c->ts_smem_set = strip_smem;
    }
    const unsigned strips = (unsigned)((nx + gpy::kTsStrip - 1) / gpy::kTsStrip);
    gpy::ts_map_strip_kernel<<<strips * (unsigned)ny, 32 * gpy::kTsStrip, strip_smem, c->stream>>>(tp);
  } else {
    gpy::ts_map_kernel<<<(unsigned)((P + 7) / 8), 256, 0, c->stream>>>(tp);
  }
  c->launches++;
  e = cudaGetLastError();
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(out, d_out.p, GPY_TS_PLANES * P * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  if (e != cudaSuccess) return done(fail(GPY_ERR_CUDA, "TS map failed: %s", cudaGetErrorString(e)));
  return done(GPY_OK);
}

int gpy_solid_angle(gpy_context_t* c, const gpy_geom_desc* g, double* out) {
  if (!c || !g || !out) return fail(GPY_ERR_INVALID, "NULL argument to gpy_solid_angle");
  CUDA_TRY(cudaSetDevice(c->device));
  const size_t P = (size_t)g->nx * g->ny;
  GPY_TRY(c->scratch_a.reserve(P * sizeof(double)));
  GeomH gh = geom_from_desc(*g);
  gpy::solid_angle_kernel<<<dim3((g->nx + 127) / 128, g->ny), 128, 0, c->stream>>>(wcs_of(gh), g->nx, g->ny,
                                                                                    (double*)c->scratch_a.p);
  c->launches++;
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(out, c->scratch_a.p, P * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  return GPY_OK;
}

int gpy_integrate_geom(gpy_context_t* c, const gpy_geom_desc* g, int32_t spatial_type, int32_t model_frame,
                       const double* sp, float* out, int32_t* oversampling) {
  if (!c || !g || !sp || !out) return fail(GPY_ERR_INVALID, "NULL argument to gpy_integrate_geom");
  if (!valid_frame(g->frame) || !valid_frame(model_frame)) return fail(GPY_ERR_UNSUPPORTED, "celestial frame code (0 galactic, 1 icrs)");
  if (spatial_type < 0 || spatial_type > 2) return fail(GPY_ERR_UNSUPPORTED, "unsupported spatial model type %d", spatial_type);
  CUDA_TRY(cudaSetDevice(c->device));
  const size_t P = (size_t)g->nx * g->ny;
  GeomH gh = geom_from_desc(*g);
  GPY_TRY(c->scratch_a.reserve(P * sizeof(double)));
  const int pitch = (g->nx + 3) / 4 * 4;
  GPY_TRY(c->scratch_b.reserve((size_t)pitch * g->ny * sizeof(float)));
  gpy::solid_angle_kernel<<<dim3((g->nx + 127) / 128, g->ny), 128, 0, c->stream>>>(wcs_of(gh), g->nx, g->ny,
                                                                                    (double*)c->scratch_a.p);
  c->launches++;
  Rect full;
  full.x0 = 0;
  full.y0 = 0;
  full.cx = g->nx;
  full.cy = g->ny;
  int f = 0;
  GPY_TRY(compute_spatial(c, gh, (const double*)c->scratch_a.p, full, spatial_type, sp, (float*)c->scratch_b.p, pitch, 0,
                          &f, nullptr, nullptr, model_frame));
  if (oversampling) *oversampling = f;
  CUDA_TRY(cudaMemcpy2DAsync(out, (size_t)g->nx * sizeof(float), c->scratch_b.p, (size_t)pitch * sizeof(float),
                             (size_t)g->nx * sizeof(float), g->ny, cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  return GPY_OK;
}

int gpy_psf_convolve(gpy_context_t* c, const float* image, int32_t ny, int32_t nx, const float* kernel,
                     int32_t n_planes, int32_t k, float* out) {
  if (!c || !image || !kernel || !out) return fail(GPY_ERR_INVALID, "NULL argument to gpy_psf_convolve");
  if (ny <= 0 || nx <= 0 || n_planes <= 0 || k <= 0 || k % 2 == 0)
    return fail(GPY_ERR_INVALID, "bad shape for gpy_psf_convolve (kernel side must be odd)");
  CUDA_TRY(cudaSetDevice(c->device));
  std::vector<float> kflip;
  std::vector<gpy::ConvPlane> planes;
  int max_h, max_kw;
  prepare_planes(kernel, n_planes, k, kflip, planes, max_h, max_kw);
  const int pitch = (nx + 3) / 4 * 4;
  DevBuf d_img, d_k, d_pl, d_out;
  int rc = GPY_OK;
  auto done = [&](int r) {
    d_img.release();
    d_k.release();
    d_pl.release();
    d_out.release();
    return r;
  };
  if ((rc = d_img.reserve((size_t)pitch * ny * sizeof(float)))) return done(rc);
  if ((rc = d_out.reserve((size_t)pitch * ny * n_planes * sizeof(float)))) return done(rc);
  cudaError_t e = cudaMemsetAsync(d_img.p, 0, (size_t)pitch * ny * sizeof(float), c->stream);
  if (e == cudaSuccess)
    e = cudaMemcpy2DAsync(d_img.p, (size_t)pitch * sizeof(float), image, (size_t)nx * sizeof(float),
                          (size_t)nx * sizeof(float), ny, cudaMemcpyHostToDevice, c->stream);
  if (e != cudaSuccess) return done(fail(GPY_ERR_CUDA, "image upload failed: %s", cudaGetErrorString(e)));
  if ((rc = upload_to(d_k, kflip.data(), kflip.size() * sizeof(float), c->stream))) return done(rc);
  if ((rc = upload_to(d_pl, planes.data(), planes.size() * sizeof(gpy::ConvPlane), c->stream))) return done(rc);
  if ((rc = launch_conv(c, (const float*)d_img.p, ny, nx, pitch, 0, (const float*)d_k.p,
                        (const gpy::ConvPlane*)d_pl.p, n_planes, max_h, max_kw, (float*)d_out.p)))
    return done(rc);
  e = cudaMemcpy2DAsync(out, (size_t)nx * sizeof(float), d_out.p, (size_t)pitch * sizeof(float),
                        (size_t)nx * sizeof(float), (size_t)ny * n_planes, cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  if (e != cudaSuccess) return done(fail(GPY_ERR_CUDA, "convolution failed: %s", cudaGetErrorString(e)));
  return done(GPY_OK);
}

int gpy_spectral_integral(gpy_context_t* c, int32_t spectral_type, const double* sp, const double* edges,
                          int32_t n_bins, const double* nodes, int32_t n_nodes, double* out) {
  if (!c || !sp || !edges || !out) return fail(GPY_ERR_INVALID, "NULL argument to gpy_spectral_integral");
  if (spectral_type < 0 || spectral_type > 2) return fail(GPY_ERR_UNSUPPORTED, "unsupported spectral model type %d", spectral_type);
  if (spectral_type != GPY_SPECTRAL_PL && (!nodes || n_nodes < 2)) return fail(GPY_ERR_INVALID, "nodes required");
  CUDA_TRY(cudaSetDevice(c->device));
  const size_t nn = nodes ? (size_t)n_bins * n_nodes : 0;
  GPY_TRY(c->scratch_a.reserve((n_bins + 1 + nn + n_bins) * sizeof(double) + sizeof(gpy::FluxJob) + 64));
  double* d_edges = (double*)c->scratch_a.p;
  double* d_nodes = d_edges + n_bins + 1;
  double* d_out = d_nodes + nn;
  gpy::FluxJob* d_job = (gpy::FluxJob*)(((uintptr_t)(d_out + n_bins) + 15) & ~(uintptr_t)15);
  gpy::FluxJob job;
  std::memset(&job, 0, sizeof(job));
  job.type = spectral_type;
  job.n_bins = n_bins;
  job.n_nodes = n_nodes;
  for (int k = 0; k < n_spectral_params(spectral_type); ++k) job.p[k] = sp[k];
  for (int k = 0; k < 5; ++k) job.pidx[k] = -1;
  job.edges = d_edges;
  job.nodes = d_nodes;
  job.out = d_out;
  CUDA_TRY(cudaMemcpyAsync(d_edges, edges, (n_bins + 1) * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  if (nn) CUDA_TRY(cudaMemcpyAsync(d_nodes, nodes, nn * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CUDA_TRY(cudaMemcpyAsync(d_job, &job, sizeof(job), cudaMemcpyHostToDevice, c->stream));
  const int seg_cap = spectral_type == GPY_SPECTRAL_PL ? 0 : std::min(n_bins * std::max(2 * n_nodes - 1, 0), 5120);
  gpy::spectral_flux_kernel<<<1, seg_cap ? 256 : 64, (size_t)seg_cap * sizeof(double), c->stream>>>(d_job, seg_cap, nullptr, nullptr);
  c->launches++;
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(out, d_out, n_bins * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  return GPY_OK;
}

}  // extern "C"
