"""The storage format of stacked fits (BASELINE.json configs[4]; SURVEY.md section 7 "Config 5 memory"): uint16
counts, bit-packed mask, one exposure upload for datasets sharing the map.  Counts are integers and the mask is
boolean either way, so the statistic must be IDENTICAL to the float32 / uint8 layout's, bit for bit, on the staged
(TMA) kernel and on the any-shape kernel; and within 1e-3 of the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _with_counts(ds, seed):
    from gammapy_b200 import Map
    from oracle import adapter

    de = adapter.evaluator_for(ds, conv="exact64")
    counts = de.fake(seed)
    ds.counts = Map.from_geom(ds._geom, data=counts, dtype=float)
    return de


@pytest.mark.parametrize("width", [(4.0, 4.0), (3.98, 2.0)])  # 200 x 200 (tensor-map staging), 199 x 100 (any-shape kernel)
def test_compact_equals_float_layout_bit_for_bit(ctx, width):
    from gammapy_b200 import configs

    ds = configs.make_c2(width=width, mask_radius=0.9)   # ragged mask, tiles of 128 px wrap rows, last tile partial
    de = _with_counts(ds, 7)
    ref = ds.stat_sum()
    assert abs(ref - de.stat_sum()) < 1e-3
    eng = ds._get_engine()
    grid = np.tile(eng.values(), (4, 1))
    col = [i for i, p in enumerate(eng.parameters) if p.name == "index"][0]
    grid[:, col] += np.linspace(0.0, 0.3, 4)
    scan_ref = eng.stat_sum(grid)

    ds.storage = "compact"
    assert ds._device is None  # the HBM mirror is rebuilt in the new layout
    got = ds.stat_sum()
    assert got == ref, (got, ref)
    dev = ds._device
    assert dev._keep["counts"].dtype.itemsize == 2
    P = ds._geom.npix[0] * ds._geom.npix[1]
    assert dev._keep["mask"].shape == (ds._geom.axes["energy"].nbin, (P + 127) // 128 * 16)
    np.testing.assert_array_equal(ds._get_engine().stat_sum(grid), scan_ref)
    # npred does not read counts / mask: unchanged
    rel = np.max(np.abs(ds.npred().data - de.npred()) / np.maximum(de.npred(), 1e-300))
    assert rel < 1e-5
    # fake() -> refresh of the uint16 counts in place
    ds.fake(random_state=3)
    c = ds.counts.data.copy()
    ds.storage = "float32"
    from gammapy_b200 import Map

    ds.counts = Map.from_geom(ds._geom, data=c, dtype=float)
    want = ds.stat_sum()
    ds.storage = "compact"
    assert ds.stat_sum() == want


def test_compact_refuses_what_it_cannot_hold(ctx):
    from gammapy_b200 import Map, configs

    ds = configs.make_c1(width=(2.0, 2.0), mask_radius=0.9)
    _with_counts(ds, 1)
    ds.storage = "compact"
    data = ds.counts.data.copy()
    data[0, 10, 10] = 70000.0
    ds.counts = Map.from_geom(ds._geom, data=data, dtype=float)
    with pytest.raises(ValueError):
        ds.stat_sum()
    data[0, 10, 10] = 2.5
    ds.counts = Map.from_geom(ds._geom, data=data, dtype=float)
    with pytest.raises(ValueError):
        ds.stat_sum()
    with pytest.raises(NotImplementedError):
        configs.make_c1(width=(2.0, 2.0)).__class__(stat_type="cash_weighted", storage="compact",
                                                    counts=ds.counts, exposure=ds.exposure)


def test_shared_exposure_is_uploaded_once_and_joint_stat(ctx):
    """Datasets of one pointing grid share the exposure Map: one HBM copy; the joint statistic of compact datasets
    equals the sum of the oracle's."""
    from gammapy_b200 import Datasets, configs

    a = configs.make_c1(name="a", width=(2.0, 2.0), mask_radius=0.9)
    b = configs.make_c1(name="b", width=(2.0, 2.0), mask_radius=0.7)
    b.exposure = a.exposure  # the same Map object
    src = [m for m in a.models if m.name == "model-simu"][0]
    b.models = [src] + [m for m in b.models if m.name != "model-simu"]
    dea, deb = _with_counts(a, 2), _with_counts(b, 3)
    a.storage = b.storage = "compact"
    joint = Datasets([a, b])
    got = joint.stat_sum()
    assert abs(got - (dea.stat_sum() + deb.stat_sum())) < 2e-3
    assert a._device._keep["exposure"].data_ptr() == b._device._keep["exposure"].data_ptr()
    assert a._device._keep["counts"].data_ptr() != b._device._keep["counts"].data_ptr()
