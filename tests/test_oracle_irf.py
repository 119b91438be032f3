"""oracle/irf.py (IRF kernel extraction at a position) on the CPU: the position-independent limit must reproduce the
kernel of the host builder the reference's ``test_npred_psf_after_edisp`` golden pins (tests/test_oracle_kat.py), a
map whose pixels all hold the same profile must give that kernel at every position, and the nearest-pixel / bilinear
rules are checked on a two-zone map."""
import numpy as np

from gammapy_b200 import configs
from oracle import adapter, irf as oirf


def _dataset():
    return configs.make_dataset(width=(4.0, 2.0), n_reco=5, n_true=8, mask_radius=None)


def test_position_independent_limit_equals_the_host_builder():
    ds = _dataset()
    want = ds.psf.get_psf_kernel(ds.exposure.geom, containment=0.999).data
    od = adapter.dataset_to_oracle(ds)
    got = oirf.psf_kernel_at(ds.psf.data, ds.psf.rad_axis.edges, None, (0.7, -0.2), od.geom.binsz,
                             tuple(od.geom.center_skydir))
    assert got.shape == want.shape
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-12)
    # a map with the same profile in every pixel: any position gives that kernel
    configs.zoned_irfs(ds, psf_scale=(1.0, 1.0), edisp_sigma=(0.15, 0.15), edisp_bias=(0.0, 0.0))
    od = adapter.dataset_to_oracle(ds)
    for pos in [(1.3, 0.4), (0.0, 0.0), (-1.9, -0.9), (5.0, 3.0)]:  # the last one is clamped to the grid
        k = oirf.psf_kernel_at(od.psf_map, od.rad_edges, od.irf_geom, pos, od.geom.binsz, tuple(od.geom.center_skydir))
        np.testing.assert_allclose(k, want, rtol=0, atol=3e-7 * want.max())  # the map is held in float32
    np.testing.assert_allclose(k.sum(axis=(1, 2)), 1.0, rtol=1e-12)


def test_two_zones_nearest_pixel_and_bilinear_profile():
    ds = _dataset()
    configs.zoned_irfs(ds)
    od = adapter.dataset_to_oracle(ds)
    ig = od.irf_geom
    assert (ig.nx, ig.ny) == (4, 2)
    # nearest IRF pixel for the edisp: pixel centres at lon = +1.5, +0.5, -0.5, -1.5 (CAR: lon decreases with x)
    m_left = oirf.edisp_kernel_at(od.edisp_map, ig, (1.4, 0.4))
    m_right = oirf.edisp_kernel_at(od.edisp_map, ig, (-0.6, -0.4))
    np.testing.assert_array_equal(m_left, od.edisp_map[:, :, 1, 0])
    np.testing.assert_array_equal(m_right, od.edisp_map[:, :, 0, 2])
    assert np.abs(m_left - m_right).max() > 0.05
    # host mirror of the product (EDispKernelMap.get_edisp_kernel) follows the same rule
    np.testing.assert_array_equal(ds.edisp.get_edisp_kernel(position=(1.4, 0.4)).pdf_matrix, m_left)
    # PSF profile: zone values at the pixel centres, their mean half-way between the two central columns
    p0 = oirf.psf_profile_at(od.psf_map, ig, 0.5, 0.0)
    p1 = oirf.psf_profile_at(od.psf_map, ig, -0.5, 0.0)
    pm = oirf.psf_profile_at(od.psf_map, ig, 0.0, 0.0)
    np.testing.assert_allclose(pm, 0.5 * (p0 + p1), rtol=1e-12)
    np.testing.assert_allclose(ds.psf.at_position((0.0, 0.0)).data, pm, rtol=1e-6)
    # containment radius grows with the zone's PSF width, kernels are normalised per plane
    r0 = oirf.containment_radius(p0, od.rad_edges, 0.68)
    r1 = oirf.containment_radius(p1, od.rad_edges, 0.68)
    assert np.all(r1 > 1.5 * r0)
    k1 = oirf.psf_kernel_at(od.psf_map, od.rad_edges, ig, (-1.0, 0.0), od.geom.binsz, tuple(od.geom.center_skydir))
    k0 = oirf.psf_kernel_at(od.psf_map, od.rad_edges, ig, (1.0, 0.0), od.geom.binsz, tuple(od.geom.center_skydir))
    assert k1.shape[-1] > k0.shape[-1]
    np.testing.assert_allclose(k1.sum(axis=(1, 2)), 1.0, rtol=1e-12)
    # product host mirror == oracle (independent code paths)
    want = ds.psf.get_psf_kernel(ds.exposure.geom, position=(-1.0, 0.0)).data
    np.testing.assert_allclose(k1, want, rtol=0, atol=3e-7 * want.max())


def test_containment_radius_against_the_reference_goldens():
    """The reference's data-free PSF-map tests (irf/psf/tests/test_map.py:429-473 ``test_psf_map_from_gauss`` /
    ``_const_sigma``; :518-536 the RecoPSFMap variant's containment values): a Gaussian PSF map built by ``from_gauss``
    has its 39.4 % containment radius at sigma within 1 %, and containment(0.1 deg) = 1 - exp(-0.5 (0.1 / sigma)^2) --
    for the host mirror (``PSFMap.containment_radius``: irf/psf/core.py:39-87) and for the oracle restatement
    (``oracle.irf.containment_radius``), which is what sizes every evaluator's kernel."""
    from gammapy_b200 import MapAxis, PSFMap

    energy_axis = MapAxis.from_nodes([1, 3, 10], name="energy_true", interp="log", unit="TeV")
    rad_axis = MapAxis.from_nodes(np.linspace(0, 1.5, 100), name="rad", unit="deg")
    for sigma in ([0.1, 0.2, 0.4], [0.1, 0.1, 0.1]):
        psf = PSFMap.from_gauss(energy_axis, rad_axis, sigma)
        assert psf.data.shape == (3, 100)
        np.testing.assert_allclose(psf.containment_radius(0.394), sigma, rtol=0.01)
        np.testing.assert_allclose(oirf.containment_radius(psf.data, rad_axis.edges, 0.394), sigma, rtol=0.01)
    psf = PSFMap.from_gauss(energy_axis, rad_axis, [0.1, 0.2, 0.3])
    np.testing.assert_allclose(psf.containment(0.1).ravel(), [0.3938, 0.1175, 0.0540], rtol=1e-2)
    # the radius that sizes the kernels, on the fine default rad axis: 99.9 % of a Gaussian lies within 3.717 sigma
    fine = PSFMap.from_gauss(energy_axis, sigma=[0.04, 0.07, 0.1])  # (the default rad axis ends at 0.66 deg)
    r999 = oirf.containment_radius(fine.data, fine.rad_axis.edges, 0.999)
    # (midpoint sums over 0.01 deg rad bins: a few per cent above the true containment, so the radius comes out low)
    np.testing.assert_allclose(r999, np.array([0.04, 0.07, 0.1]) * np.sqrt(-2 * np.log(0.001)), rtol=0.1)
    np.testing.assert_allclose(fine.containment_radius(0.999), r999, rtol=1e-12)


def test_get_psf_kernel_multiscale_reference_check():
    """datasets/tests/test_map.py:2437-2459 ``test_get_psf_kernel_multiscale`` (data-free): every plane of the kernel of
    a Gaussian PSF map with sigma 0.1 / 0.2 / 0.3 deg is cut at its own 99.9 % containment radius (the two wide ones
    hit the end of the default rad axis), and ``max_radius`` sizes the kernel map; host mirror and oracle, which must
    agree exactly in this position-independent case."""
    from gammapy_b200 import MapAxis, PSFMap, WcsGeom
    from gammapy_b200.utils import DEG, angular_separation

    energy_axis = MapAxis.from_edges(np.logspace(-1.0, 1.0, 4), unit="TeV", name="energy_true", interp="log")
    geom = WcsGeom.create(binsz=0.02, width=4.0, axes=[energy_axis])
    psf = PSFMap.from_gauss(energy_axis, sigma=[0.1, 0.2, 0.3])
    wide = psf.get_psf_kernel(geom=geom, max_radius=3.0)
    assert abs(wide.data.shape[-1] * 0.02 - 2 * 3.0) <= 0.02 + 1e-12
    kernel = psf.get_psf_kernel(geom=geom)
    k = kernel.data.shape[-1]
    off = (np.arange(k) - (k - 1) / 2) * 0.02
    sep = angular_separation(-off[None, :] * DEG, off[:, None] * DEG, 0.0, 0.0) / DEG
    for image, width in zip(kernel.data, [0.74, 1.34, 1.34]):
        assert np.all(image[sep > width] == 0) and np.any(image[sep <= width] > 0)
    np.testing.assert_allclose(kernel.data.sum(axis=(1, 2)), 1.0, rtol=1e-12)
    assert (kernel.data[0] > 0).sum() < (kernel.data[1] > 0).sum() <= (kernel.data[2] > 0).sum()
    want = oirf.psf_kernel_at(psf.data, psf.rad_axis.edges, None, (0.0, 0.0), 0.02, (0.0, 0.0))
    np.testing.assert_array_equal(kernel.data, want)
