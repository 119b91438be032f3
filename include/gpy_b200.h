/*
 * gpy_b200.h -- C ABI of libgpy_b200.so: gammapy's forward-fold likelihood hot path on B200 (sm_100a).
 *
 * One shared library, plain pointers and sizes, no torch / numpy / astropy types.  Every entry point
 * names the reference interface (path relative to /root/reference/gammapy) it stands in for.  The
 * reference is pure Python (+ one Cython file), so the binding a maintainer would add is a ctypes
 * stub -- see INTEGRATION.md.
 *
 * Conventions
 *   - units: angles deg, energies TeV, exposure m2 s, flux cm-2 s-1 TeV-1 (npred = flux*exposure*1e4,
 *     datasets/evaluator.py:346-352).
 *   - cubes are C-order (energy, lat(y), lon(x)), x fastest, as numpy / gammapy Map.data.
 *   - "DEVICE" pointers are CUDA device memory owned by the caller (e.g. torch tensors) and must
 *     outlive the dataset handle; "HOST" pointers are read during the call only.
 *   - image WCS: square pixels, cdelt = (-binsz, +binsz), CAR about the equator or TAN (gpy_projection;
 *     maps/wcs/geom.py:294-417).
 *   - every function returns 0 on success, a negative gpy_status otherwise; gpy_last_error() gives
 *     the message of the last failure on the calling thread.  There is no CPU fallback: without a
 *     CUDA device gpy_init fails with GPY_ERR_CUDA.
 *   - calls on one context are not thread-safe (the reference's caller is the single-threaded Minuit
 *     callback, modeling/iminuit.py:23-31).
 */
#ifndef GPY_B200_H
#define GPY_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GPY_ABI_VERSION 2

typedef enum {
  GPY_OK = 0,
  GPY_ERR_INVALID = -1,     /* bad argument / unsupported configuration (Python side raises ValueError) */
  GPY_ERR_CUDA = -2,        /* CUDA runtime failure, no device, launch error */
  GPY_ERR_UNSUPPORTED = -3  /* valid in the reference but outside this path's scope */
} gpy_status;

/* modeling/models/spatial.py: PointSpatialModel :558, GaussianSpatialModel :639, DiskSpatialModel :898 */
typedef enum { GPY_SPATIAL_POINT = 0, GPY_SPATIAL_GAUSS = 1, GPY_SPATIAL_DISK = 2 } gpy_spatial_type;
/* modeling/models/spectral.py: PowerLaw :1002, LogParabola :1865, ExpCutoffPowerLaw :1520 */
typedef enum { GPY_SPECTRAL_PL = 0, GPY_SPECTRAL_LP = 1, GPY_SPECTRAL_ECPL = 2 } gpy_spectral_type;

typedef struct gpy_context gpy_context_t;
typedef struct gpy_dataset gpy_dataset_t;

/* Image projections, both about the reference point (crval_lon, crval_lat) = the skydir WcsGeom.create is given
 * (maps/wcs/geom.py:370-410), with the default LONPOLE / LATPOLE: CAR (plate carree; the separable linear map when
 * crval_lat == 0, the oblique projection of wcslib otherwise) and TAN (gnomonic). */
typedef enum { GPY_PROJ_CAR = 0, GPY_PROJ_TAN = 1 } gpy_projection;
/* Celestial frames of map geometries and spatial models.  A model in another frame than the map is evaluated as the
 * reference does (SpatialModel.evaluate_geom: geom.get_coord(frame=model.frame), modeling/models/spatial.py:196-220;
 * the position goes through skycoord_to_pixel for cutouts and IRF lookups): pixel coordinates are rotated into the model
 * frame on the device, the position into the map frame on the host.  ICRS <-> Galactic as astropy defines it (frame
 * bias ICRS -> FK5 J2000 of USNO circular 179, then the IAU 1958 pole and zero point precessed to J2000). */
typedef enum { GPY_FRAME_GALACTIC = 0, GPY_FRAME_ICRS = 1 } gpy_frame;

/* WcsGeom of the counts cube (maps/wcs/geom.py:294-417) reduced to what the path reads. */
typedef struct {
  int32_t nx, ny;           /* image size in pixels (lon, lat) */
  int32_t n_true, n_reco;   /* energy_true / energy bins */
  double binsz;             /* deg */
  double crval_lon;         /* deg */
  double crpix_x, crpix_y;  /* FITS 1-based reference pixel */
  double crval_lat;         /* deg; CAR: inside (-90, 90) */
  int32_t proj;             /* gpy_projection */
  int32_t frame;            /* gpy_frame */
} gpy_geom_desc;

/* The MapDataset attributes the path consumes (datasets/map.py:410; §8 a4-a6, a17 of SURVEY.md). */
typedef struct {
  gpy_geom_desc geom;
  const double* energy_true_edges;   /* HOST [n_true+1] */
  const double* energy_reco_center;  /* HOST [n_reco]; MapAxis.center of the reco axis (FoVBackgroundModel.evaluate_geom) */
  const float* exposure;             /* DEVICE [n_true, ny, nx] */
  const float* background;           /* DEVICE [n_reco, ny, nx] or NULL */
  const float* counts;               /* DEVICE [n_reco, ny, nx] or NULL (required for stat) */
  const uint8_t* mask;               /* DEVICE [n_reco, ny, nx] or NULL: mask_safe*mask_fit (datasets/core.py:73-81) */
  const uint8_t* mask_image;         /* HOST [ny, nx] or NULL (= all true): MapDataset.mask_image (datasets/map.py:1040-1048) */
  const float* psf_kernel;           /* HOST [n_true, psf_k, psf_k] float32, PSFKernel.data planes normalised (irf/psf/kernel.py:70-80), or NULL */
  int32_t psf_k;                     /* odd kernel side */
  const double* edisp;               /* HOST [n_true, n_reco] EDispKernel.pdf_matrix (irf/edisp/kernel.py:71-77) */
  const double* edisp_diagonal;      /* HOST [n_true, n_reco] diagonal response (evaluator.py:245-250) or NULL */
  const double* nodes;               /* HOST [n_true, n_nodes] geomspace nodes of integrate_spectrum (spectral.py:132-134) */
  int32_t n_nodes;                   /* int(max(100*log10(e_hi/e_lo), 2)) evaluated by the caller with numpy */
  const float* weight;               /* DEVICE [n_reco, ny, nx] or NULL: float mask of stat_type "cash_weighted"
                                        (WeightedCashFitStatistic, stats/fit_statistics.py:301-329); bins with
                                        weight 0 are excluded, `mask` must then be (weight != 0) */
  int32_t compact;                   /* 0: counts float32, mask uint8 (above).  1: the storage format of stacked fits
                                        with many datasets per GPU (BASELINE.json configs[4]): `counts` points to
                                        DEVICE uint16 [n_reco, ny, nx] (counts above 65535 cannot be stored: the
                                        caller must refuse them), `mask` to DEVICE bytes [n_reco, mask_row_bytes]
                                        holding one bit per bin, bit (p & 7) of byte (p >> 3) for pixel p = y nx + x,
                                        rows zero padded.  Same statistic: counts are integers either way. */
  int64_t mask_row_bytes;            /* compact: bytes per mask row, a multiple of 16, >= ceil(ny nx / 8) */
} gpy_dataset_desc;

/* One SkyModel: where its parameters sit in the parameter vector handed to gpy_stat_sum / gpy_npred.
 * Parameter order follows the reference classes:
 *   point: lon_0, lat_0            gauss: lon_0, lat_0, sigma, e, phi     disk: lon_0, lat_0, r_0, e, phi, edge_width
 *   pl: index, amplitude, reference      lp: amplitude, reference, alpha, beta
 *   ecpl: index, amplitude, reference, lambda_, alpha
 */
typedef struct {
  int32_t spatial_type;     /* gpy_spatial_type */
  int32_t spectral_type;    /* gpy_spectral_type */
  int32_t spatial_par[6];   /* column of each spatial parameter in params[b, :] (unused entries -1) */
  int32_t spectral_par[5];  /* column of each spectral parameter */
  int32_t apply_psf;        /* SkyModel.apply_irf["psf"]   (modeling/models/cube.py) */
  int32_t apply_edisp;      /* SkyModel.apply_irf["edisp"] : 0 -> diagonal response (evaluator.py:371-377) */
  int32_t frame;            /* gpy_frame of the spatial model (lon_0, lat_0, phi are in this frame) */
} gpy_source_desc;

/* FoVBackgroundModel with PowerLawNormSpectralModel (modeling/models/cube.py:622-756, spectral.py:1134-1162). */
typedef struct {
  int32_t enabled;          /* 0: npred_background = background (datasets/map.py:812-813) */
  int32_t par_norm, par_tilt, par_reference;
} gpy_background_desc;

/* Evaluator state of one source (MapEvaluator: datasets/evaluator.py:174-243), for inspection / tests. */
typedef struct {
  int32_t initialised;      /* update() has run */
  int32_t contributes;      /* SkyModel.contributes(mask, margin=psf_width) */
  int32_t x0, y0, cx, cy;   /* exposure cutout in parent pixels (parent slices of _cutout_view) */
  int32_t oversampling;     /* factor used by the last integrate_geom */
  int32_t n_spatial_evals;  /* how often spatial + PSF were recomputed (cache behaviour, evaluator.py:282-286) */
} gpy_source_state;

const char* gpy_last_error(void);
int gpy_abi_version(void);

/* Create a context on CUDA device `device`; `stream` is a cudaStream_t (NULL: the library makes its own). */
int gpy_init(int device, void* stream, gpy_context_t** ctx);
int gpy_finalize(gpy_context_t* ctx);
/* Number of this library's kernel launches since gpy_init (bench.py "gpu_launches"). */
int64_t gpy_launch_count(const gpy_context_t* ctx);

/* Measurement aid (bench.py roofline): with enable != 0 every evaluation brackets its fold + Cash kernel with CUDA
 * events on the library's stream; gpy_kernel_time returns the summed duration [ms] and the number of launches
 * since the last call with enable != 0, after waiting for the stream.  Costs one host wait per evaluation while on. */
int gpy_kernel_timing(gpy_context_t* ctx, int32_t enable);
int gpy_kernel_time(gpy_context_t* ctx, double* total_ms, int64_t* launches);

/* MapDataset.__init__ side: register the resident cubes, upload the static IRF kernels. */
int gpy_dataset_create(gpy_context_t* ctx, const gpy_dataset_desc* desc, gpy_dataset_t** ds);
int gpy_dataset_destroy(gpy_dataset_t* ds);
/* Replace the counts cube pointer (MapDataset.fake overwrites counts, datasets/map.py:1538-1554); same element type as
 * at creation (float32, or uint16 for a compact dataset). */
int gpy_dataset_set_counts(gpy_dataset_t* ds, const float* counts_device);
/* MapDataset.models setter (datasets/map.py:674-694): one evaluator per SkyModel, caches reset. */
int gpy_dataset_set_models(gpy_dataset_t* ds, const gpy_source_desc* sources, int32_t n_sources,
                           const gpy_background_desc* background);
int gpy_dataset_source_state(const gpy_dataset_t* ds, int32_t source, gpy_source_state* out);

/* Position-dependent IRFs: PSFMap / EDispKernelMap with a spatial grid (irf/psf/map.py, irf/edisp/map.py).  Once set,
 * every evaluator looks its own IRFs up at the source position whenever MapEvaluator.update runs
 * (datasets/evaluator.py:174-243, i.e. at the first evaluation and when the source has drifted further than
 * evaluation_radius + 0.1 deg from the position of the last lookup, :465-481):
 *   - PSFMap.get_psf_kernel(position, geom, containment=0.999) (irf/psf/map.py:249-337): radial profile interpolated
 *     bilinearly in the IRF grid at the position (clamped to the grid), containment radii from the cumulative profile
 *     (irf/psf/core.py:39-87), per-plane cut and <= 4 x oversampling (irf/psf/map.py:23-37), block sum, plane
 *     normalisation (irf/psf/kernel.py:70-80) -- profile and radii on the host, the kernel images on the device;
 *   - EDispKernelMap.get_edisp_kernel(position) (irf/edisp/map.py:369-401): the matrix of the nearest IRF pixel.
 *     Any number of matrices per dataset (one per source in the limit); the staged fold kernels serve launches whose
 *     matrices fit their shared memory (one for the 64-pixel kernel, two or three for the 128-pixel one), beyond that
 *     the any-shape kernel runs, which works through the matrices of each tile's own sources one at a time.
 * psf_map / edisp_map are HOST arrays, copied.  The psf_kernel / edisp of gpy_dataset_desc are then unused for the
 * respective IRF (edisp_diagonal still serves apply_irf["edisp"] = False).  A batched call (n_vec > 1) whose vector
 * would trigger such a lookup fails with GPY_ERR_UNSUPPORTED: evaluate that point as a single vector first. */
typedef struct {
  int32_t nx, ny;            /* IRF image grid (same conventions as gpy_geom_desc) */
  double binsz, crval_lon, crpix_x, crpix_y;
  double crval_lat;          /* reference latitude, deg */
  int32_t proj;              /* gpy_projection */
  int32_t n_rad;             /* rad bins of the PSF map */
  const double* rad_edges;   /* HOST [n_rad + 1] deg, increasing */
  const float* psf_map;      /* HOST [n_true, n_rad, ny, nx] sr-1 at the rad bin centres, or NULL */
  const double* edisp_map;   /* HOST [n_true, n_reco, ny, nx] EDispKernelMap data, or NULL */
} gpy_irf_maps_desc;
int gpy_dataset_set_irf_maps(gpy_dataset_t* ds, const gpy_irf_maps_desc* desc);
/* Inspection (tests): the kernel an evaluator currently uses.  *k = its side; kernel HOST [n_true, k, k] or NULL
 * (capacity in floats); *edisp_pixel = iy * nx + ix of the IRF pixel of its edisp matrix (-1 without an edisp map);
 * *n_updates = how often its IRFs were looked up.  The two last pointers may be NULL. */
int gpy_dataset_source_psf_kernel(const gpy_dataset_t* ds, int32_t source, float* kernel, int32_t capacity, int32_t* k,
                                  int32_t* edisp_pixel, int32_t* n_updates);

/* Datasets.stat_sum (datasets/core.py:264-277) for a batch of parameter vectors.
 *   params  HOST [n_vec, n_par] parameter VALUES (not factors); shared by all datasets.
 *   stat    HOST [n_vec]        sum over datasets (summed in dataset order, as the Python loop does)
 *   stat_per_dataset HOST [n_vec, n_datasets] or NULL
 * n_vec == 1 advances the evaluators' cache state exactly like one reference evaluation
 * (MapEvaluator.needs_update / update / parameters_spatial_changed).  n_vec > 1 evaluates each
 * vector from the current state without mutating it (profile scans, finite-difference gradients);
 * work items (tile, vector) whose inputs equal those of vector 0 on the same tile are evaluated once,
 * so put the base point of a scan or gradient first.
 */
int gpy_stat_sum(gpy_context_t* ctx, gpy_dataset_t* const* datasets, int32_t n_datasets,
                 const double* params, int32_t n_vec, int32_t n_par, double* stat,
                 double* stat_per_dataset);

/* Same evaluation, leaving the per-(vector, dataset) stats on the device:
 *   stat_device DEVICE [n_vec, n_datasets] float64.  No host synchronisation for n_vec == 1: the
 *   caller owns the stream order (e.g. an NCCL all-reduce of the buffer on the same stream).  A call
 *   with n_vec > 1 waits once, for the 4-byte length of its worklist, before it launches the fold. */
int gpy_stat_sum_device(gpy_context_t* ctx, gpy_dataset_t* const* datasets, int32_t n_datasets,
                        const double* params, int32_t n_vec, int32_t n_par, double* stat_device);

/* Joint fits over ranks, one process per GPU -- replaces the reference's sum over Ray actors
 * (datasets/actors.py:84-87 `sum(ray.get([actor.stat_sum(...)]))`, 168-173) by a one-shot all-reduce over NVLink peer
 * memory that runs as the last kernel of the evaluation (no NCCL launch, no host involvement).
 *
 * gpy_peer_reduce_bytes: size of the buffer every rank must provide for `world` ranks and batches of up to `capacity`
 *   parameter vectors.
 * gpy_peer_reduce_init: peer_buffers[r] = this process' device mapping of rank r's buffer (symmetric memory: CUDA
 *   IPC / VMM handles, e.g. torch.distributed._symmetric_memory buffer_ptrs), all zero-filled and all ranks
 *   synchronised before the first evaluation.  Every rank must then issue the same sequence of
 *   gpy_stat_sum_joint_device calls.
 * gpy_stat_sum_joint_device: as gpy_stat_sum_device, then joint_device DEVICE [n_vec] float64 = sum over ranks of
 *   (sum over this rank's datasets, in dataset order), ranks added in rank order: bit-identical on every rank.  A rank
 *   that waits longer than 5 s for a peer writes NaN instead of hanging the device.
 * gpy_stat_sum_joint: the same, with the result copied to joint HOST [n_vec] and the stream synchronised -- the
 *   sharded counterpart of gpy_stat_sum (ShardedDatasets.stat_sum: what Fit calls on every rank). */
int64_t gpy_peer_reduce_bytes(int32_t world, int32_t capacity);
int gpy_peer_reduce_init(gpy_context_t* ctx, void* const* peer_buffers, int32_t rank, int32_t world, int32_t capacity);
int gpy_stat_sum_joint_device(gpy_context_t* ctx, gpy_dataset_t* const* datasets, int32_t n_datasets,
                              const double* params, int32_t n_vec, int32_t n_par, double* joint_device);
int gpy_stat_sum_joint(gpy_context_t* ctx, gpy_dataset_t* const* datasets, int32_t n_datasets, const double* params,
                       int32_t n_vec, int32_t n_par, double* joint);

/* MapDataset.npred (datasets/map.py:775-788) / npred_signal (:824-880) for one parameter vector.
 *   source_enable HOST [n_sources] or NULL (all): npred_signal(model_names=...)
 *   include_background: add npred_background (:790-813) and clip negatives to 0 (:787)
 *   npred_device DEVICE [n_reco, ny, nx] float64
 */
int gpy_npred(gpy_context_t* ctx, gpy_dataset_t* ds, const double* params, int32_t n_par,
              const uint8_t* source_enable, int32_t include_background, double* npred_device);

/* Compiled-stat registry entries (utils/compilation.py:35-87; stats/fit_statistics_cython.pyx:17-85):
 * HOST float64 arrays in, HOST scalar out.  cash_sum_compiled / weighted_cash_sum_compiled. */
int gpy_cash_sum(gpy_context_t* ctx, const double* counts, const double* npred, int64_t n, double* out);
int gpy_weighted_cash_sum(gpy_context_t* ctx, const double* counts, const double* npred,
                          const double* weight, int64_t n, double* out);

/* WStatFitStatistic.stat_sum_dataset (stats/fit_statistics.py:332-361) = sum over the masked bins of
 * np.nan_to_num(wstat(n_on, n_off, alpha, mu_sig)) (:142-252: profile-likelihood background :221-236, goodness-of-fit
 * terms :239-252) -- the statistic of MapDatasetOnOff (datasets/map.py:2554).  All arrays DEVICE, length n:
 *   n_on, n_off, alpha  float32 (float64_inputs == 0) or float64;  mu_sig float64 (npred_signal, e.g. the output of
 *   gpy_npred without background);  mask uint8 or NULL;  stat_per_bin DEVICE float64 [n] or NULL (stat_array_dataset);
 *   out HOST: the sum.  Waits for the result. */
int gpy_wstat_sum_device(gpy_context_t* ctx, const void* n_on, const void* n_off, const void* alpha, int32_t float64_inputs,
                         const double* mu_sig, const uint8_t* mask, int64_t n, double* stat_per_bin, double* out);

/* f_cash_root_compiled(x, counts, background, model) and norm_bounds_compiled(counts, background, model)
 * (stats/fit_statistics_cython.pyx:90-157): HOST float64 arrays in, HOST out (1 and 3 doubles:
 * b_min, b_max, -sn_min_total). */
int gpy_f_cash_root(gpy_context_t* ctx, double x, const double* counts, const double* background,
                    const double* model, int64_t n, double* out);
int gpy_norm_bounds(gpy_context_t* ctx, const double* counts, const double* background,
                    const double* model, int64_t n, double* out3);

/* TS map: the per-pixel one-parameter fit TSMapEstimator.estimate_flux_map distributes over processes
 * (estimators/map/ts.py:516-531 -> _ts_value :1098-1153 -> BrentqFluxEstimator.run :1053-1095 with
 * estimate_best_fit :819-870 on SimpleMapDataset.from_arrays :761-783), default selection, ts_threshold None.
 *   counts / background / exposure  HOST [n_planes, ny, nx] float64 (reco-energy planes; several datasets of the
 *                                   same image geometry are concatenated along the plane axis, as _ts_value does)
 *   kernel                          HOST [n_planes, ky, kx] float64, ky and kx odd
 *   mask                            HOST [ny, nx] uint8 or NULL: positions to fit (_estimate_mask_2d); others get NaN
 *   rtol, max_niter                 scipy.optimize.brentq arguments (xtol is scipy's default 2e-12)
 *   out                             HOST [GPY_TS_PLANES, ny, nx] float64: norm, norm_err, niter, ts, stat, stat_null,
 *                                   success, npred, npred_excess
 * Cutout bins outside the map are absent, which is what the reference's zero padding amounts to (:772-777). */
#define GPY_TS_PLANES 9
int gpy_ts_map(gpy_context_t* ctx, const double* counts, const double* background, const double* exposure,
               int32_t n_planes, int32_t ny, int32_t nx, const double* kernel, int32_t ky, int32_t kx,
               const uint8_t* mask, double rtol, int32_t max_niter, double n_sigma, double* out);

/* Single-stage entry points (the parity tests exercise each kernel through these).
 * SpatialModel.integrate_geom on an image geometry (modeling/models/spatial.py:222-309, 609-631):
 *   spatial_params HOST in the class order above; out HOST [ny, nx] float32 (the value the PSF
 *   convolution consumes, maps/wcs/ndmap.py:984); oversampling: factor used. */
int gpy_integrate_geom(gpy_context_t* ctx, const gpy_geom_desc* geom, int32_t spatial_type, int32_t model_frame,
                       const double* spatial_params, float* out, int32_t* oversampling);
/* WcsNDMap.convolve(PSFKernel) of one image with n_planes kernels, mode "same" (maps/wcs/ndmap.py:920-1010):
 *   image HOST [ny, nx] float32, kernel HOST [n_planes, k, k] float32, out HOST [n_planes, ny, nx] float32. */
int gpy_psf_convolve(gpy_context_t* ctx, const float* image, int32_t ny, int32_t nx, const float* kernel,
                     int32_t n_planes, int32_t k, float* out);
/* SpectralModel.integral over bins (modeling/models/spectral.py:353-371, 103-164, 1039-1067):
 *   edges HOST [n_bins+1], nodes HOST [n_bins, n_nodes], out HOST [n_bins]. */
int gpy_spectral_integral(gpy_context_t* ctx, int32_t spectral_type, const double* spectral_params,
                          const double* edges, int32_t n_bins, const double* nodes, int32_t n_nodes,
                          double* out);
/* WcsGeom.solid_angle (maps/wcs/geom.py:803-854): out HOST [ny, nx] float64 sr. */
int gpy_solid_angle(gpy_context_t* ctx, const gpy_geom_desc* geom, double* out);

#ifdef __cplusplus
}
#endif
#endif /* GPY_B200_H */
