"""``MapDataset.stack`` / ``to_masked`` / ``from_geoms`` / ``Datasets.stack_reduce`` (datasets/map.py:883-945, 1104-1210,
1958-1986; datasets/core.py:530-569) against the numbers of the reference's data-free ``test_stack``
(datasets/tests/test_map.py:1129-1235): host arithmetic, no GPU (datasets without a background model)."""
import numpy as np
import pytest
from numpy.testing import assert_allclose

from gammapy_b200 import Datasets, EDispKernelMap, Map, MapAxis, MapDataset, PSFMap, WcsGeom


def _pair():
    axis = MapAxis.from_energy_bounds("0.1 TeV", "10 TeV", nbin=3)
    axis_etrue = MapAxis.from_energy_bounds("0.1 TeV", "10 TeV", nbin=5, name="energy_true")
    geom = WcsGeom.create(skydir=(266.40498829, -28.93617776), binsz=0.05, width=(2, 2), frame="icrs", axes=[axis])
    geom_etrue = geom.to_image().to_cube([axis_etrue])
    edisp = EDispKernelMap.from_diagonal_response(energy_axis=axis, energy_axis_true=axis_etrue)
    out = []
    for k, (bkg, rows) in enumerate(((0.2, [(0, 5, 10)]), (0.1, [(0, 5, 10), (1, 10, 15)]))):
        mask = np.ones(geom.data_shape, dtype=bool)
        for e, lo, hi in rows:
            mask[e][lo:hi] = False
        out.append(MapDataset(
            counts=Map.from_geom(geom, data=np.ones(geom.data_shape)),
            background=Map.from_geom(geom, data=np.full(geom.data_shape, bkg)),
            exposure=Map.from_geom(geom_etrue, data=np.full(geom_etrue.data_shape, 1e7), unit="m2 s"),
            mask_safe=Map.from_geom(geom, data=mask.copy(), dtype=bool), mask_fit=Map.from_geom(geom, data=mask.copy(), dtype=bool),
            edisp=edisp, name=f"dataset-{k + 1}"))
    return out


def test_stack_reference_numbers():
    dataset1, dataset2 = _pair()
    assert dataset1._geom.data_shape == (3, 40, 40)  # an oblique CAR field about the galactic centre, in ICRS
    stacked = MapDataset.from_geoms(**dataset1.geoms)
    stacked.stack(dataset1)
    stacked.stack(dataset2)
    assert_allclose(stacked.background.data.sum(), 1360, rtol=1e-5)
    assert_allclose(stacked.counts.data.sum(), 9000, rtol=1e-5)
    assert stacked.mask_safe.data.sum() == 4600 and stacked.mask_fit.data.sum() == 4600
    assert_allclose(stacked.exposure.data.sum(), 1.6e11)
    np.testing.assert_array_equal(stacked.edisp.get_edisp_kernel().pdf_matrix, dataset1.edisp.get_edisp_kernel().pdf_matrix)
    # without safe masks everything enters, every time
    plain = MapDataset(counts=dataset1.counts, background=dataset1.background)
    total = MapDataset.from_geoms(**plain.geoms)
    for _ in range(3):
        total.stack(plain)
    assert_allclose(total.background.data.sum(), 2880.0, rtol=1e-5)
    assert_allclose(total.counts.data.sum(), 14400.0, rtol=1e-5)
    # an all-false safe mask contributes nothing
    empty = MapDataset(counts=dataset1.counts, background=dataset1.background,
                       mask_safe=Map.from_geom(dataset1._geom, dtype=bool))
    total.stack(empty)
    assert_allclose(total.counts.data.sum(), 14400.0, rtol=1e-5)


def test_to_masked_and_stack_reduce():
    dataset1, dataset2 = _pair()
    masked = dataset1.to_masked(name="masked")
    safe = np.asarray(dataset1.mask_safe.data, bool)
    assert masked.name == "masked" and masked.counts.data.sum() == safe.sum()
    assert np.all(masked.counts.data[~safe] == 0) and np.all(masked.background.data[~safe] == 0)
    assert_allclose(masked.background.data[safe], 0.2)
    reduced = Datasets([dataset1, dataset2]).stack_reduce(name="stacked")
    assert reduced.name == "stacked"
    assert_allclose(reduced.counts.data.sum(), 9000)
    assert_allclose(reduced.background.data.sum(), 1360)
    assert reduced.mask_safe.data.sum() == 4600
    assert dataset1.counts.data.sum() == 4800  # the inputs are untouched


def test_irfs_are_averaged_with_exposure_weights():
    dataset1, dataset2 = _pair()
    e_true = dataset1.exposure.geom.axes["energy_true"]
    dataset1.psf = PSFMap.from_gauss(e_true, sigma=0.1)
    dataset2.psf = PSFMap.from_gauss(e_true, sigma=0.3)
    dataset2.exposure = Map.from_geom(dataset2.exposure.geom, data=np.asarray(dataset2.exposure.data) * 3.0, unit="m2 s")
    e_reco = dataset1._geom.axes["energy"]
    dataset2.edisp = EDispKernelMap.from_gauss(energy_axis=e_reco, energy_axis_true=e_true, sigma=0.2, bias=0.0)
    stacked = Datasets([dataset1, dataset2]).stack_reduce()
    assert_allclose(stacked.psf.data, 0.25 * dataset1.psf.data + 0.75 * dataset2.psf.data, rtol=1e-12)
    k1, k2 = dataset1.edisp.get_edisp_kernel().pdf_matrix, dataset2.edisp.get_edisp_kernel().pdf_matrix
    assert_allclose(stacked.edisp.get_edisp_kernel().pdf_matrix, 0.25 * k1 + 0.75 * k2, rtol=1e-12)
    assert_allclose(stacked.exposure.data.sum(), 1e7 * 5 * 1600 * 4)


def test_to_masked_reference_number():
    """datasets/tests/test_map.py:2413-2421 ``test_to_masked``: 200 unit counts, 30 of them outside the safe mask."""
    axis = MapAxis.from_energy_bounds(1, 10, 2, unit="TeV")
    geom = WcsGeom.create(npix=(10, 10), binsz=0.05, axes=[axis])
    mask = np.ones(geom.data_shape, dtype=bool)
    mask[0][5:8] = False
    dataset = MapDataset(counts=Map.from_geom(geom, data=np.ones(geom.data_shape)),
                         mask_safe=Map.from_geom(geom, data=mask, dtype=bool))
    assert_allclose(dataset.to_masked().counts.data.sum(), 170)


def _on_off(alpha, seed, mask_rows=()):
    from gammapy_b200 import MapDatasetOnOff

    axis = MapAxis.from_energy_bounds(1, 10, 2, unit="TeV")
    geom = WcsGeom.create(npix=(10, 10), binsz=0.05, axes=[axis])
    rng = np.random.RandomState(seed)
    mask = np.ones(geom.data_shape, dtype=bool)
    for e, lo, hi in mask_rows:
        mask[e][lo:hi] = False
    return MapDatasetOnOff(counts=Map.from_geom(geom, data=rng.poisson(3.0, geom.data_shape).astype(float)),
                           counts_off=Map.from_geom(geom, data=rng.poisson(2.0, geom.data_shape).astype(float)),
                           acceptance=Map.from_geom(geom, data=np.ones(geom.data_shape)),
                           acceptance_off=Map.from_geom(geom, data=np.full(geom.data_shape, 1.0 / alpha)),
                           mask_safe=Map.from_geom(geom, data=mask, dtype=bool), name=f"onoff-{seed}")


def test_on_off_stacking():
    """``MapDatasetOnOff.stack`` (datasets/map.py:2936-3006): counts and ON acceptance add up under the safe mask, the
    stacked alpha is the OFF-count weighted mean of the alphas; ``to_masked`` as in the reference's ``test_to_masked``
    (datasets/tests/test_map.py:2423-2435)."""
    a, b = _on_off(0.2, 1), _on_off(0.5, 2, mask_rows=[(0, 5, 8)])
    safe_b = np.asarray(b.mask_safe.data, bool)
    stacked = a.to_masked(name="stacked")
    assert type(stacked) is type(a) and stacked.counts.data.sum() == a.counts.data.sum()
    stacked.stack(b)
    assert_allclose(stacked.counts.data, a.counts.data + b.counts.data * safe_b)
    assert_allclose(stacked.counts_off.data, a.counts_off.data + b.counts_off.data * safe_b)
    assert_allclose(stacked.acceptance.data, 1.0 + safe_b)
    off_a, off_b = a.counts_off.data, b.counts_off.data * safe_b
    seen = (off_a + off_b) > 0
    want = (0.2 * off_a + 0.5 * off_b)[seen] / (off_a + off_b)[seen] / (1.0 + safe_b)[seen]
    assert_allclose((1.0 / stacked.acceptance_off.data)[seen], want, rtol=1e-12)   # 1 / a_off = alpha / a_on
    mean_alpha = (0.2 * off_a + 0.5 * off_b).sum() / (off_a + off_b).sum()
    assert_allclose(stacked.alpha.data[~seen], mean_alpha, rtol=1e-12)
    assert stacked.mask_safe.data.all()
    # stacking a dataset onto a copy of itself doubles the counts and keeps alpha
    twice = a.to_masked()
    twice.stack(a)
    assert_allclose(twice.counts_off.data, 2 * a.counts_off.data)
    assert_allclose(twice.alpha.data[a.counts_off.data > 0], 0.2, rtol=1e-12)
    # the reference's to_masked number: 200 unit counts, 30 masked
    ones = _on_off(0.1, 3, mask_rows=[(0, 5, 8)])
    ones.counts = Map.from_geom(ones._geom, data=np.ones(ones._geom.data_shape))
    assert_allclose(ones.to_masked().counts.data.sum(), 170)
    with pytest.raises(TypeError):
        a.stack(MapDataset(counts=a.counts))


def test_dataset_cutout():
    """``MapDataset.cutout`` (datasets/map.py:2040-2094; the aligned-geometry property of the reference's
    ``test_dataset_cutout_aligned``, datasets/tests/test_map.py:1744-1756): maps sliced with the Cutout2D rules, models
    kept, and -- for a source whose evaluation region lies inside the cutout -- the same npred as the parent there."""
    from gammapy_b200 import configs
    from oracle import adapter

    ds = configs.make_c1(width=(3.0, 3.0), mask_radius=1.2)   # Gaussian source near the centre + background model
    rng = np.random.RandomState(0)
    ds.counts = Map.from_geom(ds._geom, data=rng.poisson(1.0, ds._geom.data_shape).astype(float))
    cut = ds.cutout(position=(0.1, -0.05), width=(2.4, 2.0), name="sub")
    ys, xs = ds._geom.cutout_slices_for((0.1, -0.05), (2.4, 2.0))
    assert cut.name == "sub" and cut._geom.data_shape[1:] == (ys.stop - ys.start, xs.stop - xs.start)
    assert ds._geom.is_aligned(cut._geom) and ds.exposure.geom.is_aligned(cut.exposure.geom)
    for attr in ("counts", "exposure", "background", "mask_safe", "mask_fit"):
        np.testing.assert_array_equal(getattr(cut, attr).data, getattr(ds, attr).data[..., ys, xs])
    # (the background model names the parent dataset: a cutout under another name keeps the sources only, as in the reference)
    assert cut.mask_safe.data.dtype == bool and [m.name for m in cut.models] == [m.name for m in ds.models if m.name != "c1-bkg"]
    assert [m.name for m in ds.cutout(position=(0.1, -0.05), width=(2.4, 2.0), name=ds.name).models] == [m.name for m in ds.models]
    cut.models = [m.copy(name=m.name) if m.name != "c1-bkg" else type(m)(dataset_name="sub") for m in ds.models]
    a = adapter.evaluator_for(ds, conv="exact64").npred_signal()[..., ys, xs]
    b = adapter.evaluator_for(cut, conv="exact64").npred_signal()
    inner = (slice(None), slice(22, -22), slice(22, -22))  # further from the cutout's edge than the PSF kernel reaches (21 px)
    np.testing.assert_allclose(b[inner], a[inner], rtol=1e-9, atol=1e-12 * a.max())


def test_dataset_copy():
    """``Dataset.copy`` (datasets/core.py:96-114): the data copied, a new name, no models."""
    dataset1, _ = _pair()
    from gammapy_b200.modeling.models import FoVBackgroundModel

    dataset1.models = [FoVBackgroundModel(dataset_name=dataset1.name)]
    clone = dataset1.copy(name="clone")
    assert clone.name == "clone" and clone.models is None and dataset1.models is not None
    assert clone.counts is not dataset1.counts and clone.counts.data is not dataset1.counts.data
    for attr in ("counts", "exposure", "background", "mask_safe", "mask_fit"):
        np.testing.assert_array_equal(getattr(clone, attr).data, getattr(dataset1, attr).data)
    clone.counts.data[0, 0, 0] = 99.0
    assert dataset1.counts.data[0, 0, 0] == 1.0
    assert dataset1.copy().name != dataset1.name
    on_off = _on_off(0.2, 5)
    twin = on_off.copy(name="twin")
    assert type(twin) is type(on_off) and twin.stat_type == "wstat"
    np.testing.assert_array_equal(twin.counts_off.data, on_off.counts_off.data)
    np.testing.assert_array_equal(twin.alpha.data, on_off.alpha.data)
