"""SURVEY.md 8f N4 (ingestion half): the from-scratch FITS layer and the GADF map-dataset layout (gammapy_b200/fitsio.py).
Round trips and byte-level structure; the reference's astropy-based reader cannot be run here (interoperability is by
construction: HDU names, BANDS tables and WCS keywords as datasets/map.py:1556-1700 / maps/axes.py:1963-2080 write them)."""
import numpy as np
import pytest

from gammapy_b200 import MapDataset, configs, fitsio


def test_fits_blocks_cards_and_types(tmp_path):
    path = str(tmp_path / "t.fits")
    img = np.arange(24, dtype=np.float32).reshape(2, 3, 4)
    table = {"CHANNEL": np.arange(3, dtype=np.int64), "E_MIN": np.array([1.0, 2.0, 4.0]), "FLAG": np.array([True, False, True]),
             "NAME": np.array(["a", "bc", "def"])}
    fitsio.write_hdus(path, [fitsio.HDU("PRIMARY", None, {"NAME": "x"}), fitsio.HDU("IMG", img, {"CRVAL1": -0.5, "BUNIT": "m2 s"}),
                             fitsio.HDU("TAB", table, {"AXCOLS1": "E_MIN,E_MAX"}, {"E_MIN": "TeV"})])
    raw = open(path, "rb").read()
    assert len(raw) % 2880 == 0 and raw[:30] == b"SIMPLE  =                    T"
    assert b"XTENSION= 'IMAGE   '" in raw and b"XTENSION= 'BINTABLE'" in raw and b"TUNIT2  = 'TeV     '" in raw
    with pytest.raises(OSError):
        fitsio.write_hdus(path, [fitsio.HDU("PRIMARY")])
    p, i, t = fitsio.read_hdus(path)
    assert p.header["NAME"] == "x" and p.data is None
    np.testing.assert_array_equal(i.data, img)
    assert i.data.dtype == np.float32 and i.header["CRVAL1"] == -0.5 and i.header["BUNIT"] == "m2 s"
    np.testing.assert_array_equal(t.data["CHANNEL"], [0, 1, 2])
    np.testing.assert_array_equal(t.data["E_MIN"], [1.0, 2.0, 4.0])
    np.testing.assert_array_equal(t.data["FLAG"], [True, False, True])
    assert list(t.data["NAME"]) == ["a", "bc", "def"] and t.units["E_MIN"] == "TeV" and t.header["AXCOLS1"] == "E_MIN,E_MAX"
    # big-endian payload: the first float32 of the image is 0.0, the second 1.0 = 3f800000
    off = raw.index(b"XTENSION= 'IMAGE   '")
    off += 2880 * ((raw[off:].index(b"END" + b" " * 77) // 2880) + 1)
    assert raw[off + 4:off + 8] == bytes.fromhex("3f800000")


def _assert_same_map(a, b):
    assert a.geom == b.geom
    assert [ax.name for ax in a.geom.axes] == [ax.name for ax in b.geom.axes]
    for x, y in zip(a.geom.axes, b.geom.axes):
        np.testing.assert_array_equal(x.edges, y.edges)
        assert x.interp == y.interp
    np.testing.assert_array_equal(np.asarray(a.data), np.asarray(b.data))


@pytest.mark.parametrize("zoned", [False, True])
def test_map_dataset_round_trip(tmp_path, zoned):
    ds = configs.make_dataset(width=(3.0, 1.6), n_reco=5, n_true=8, mask_radius=0.7, name="obs-7")
    if zoned:
        configs.zoned_irfs(ds)
    rng = np.random.RandomState(0)
    from gammapy_b200 import Map

    ds.counts = Map.from_geom(ds._geom, data=rng.poisson(2.0, ds._geom.data_shape).astype(np.float32))
    path = str(tmp_path / "ds.fits")
    ds.write(path)
    names = [h.name for h in fitsio.read_hdus(path)]
    assert names == ["PRIMARY", "COUNTS", "COUNTS_BANDS", "EXPOSURE", "EXPOSURE_BANDS", "BACKGROUND", "BACKGROUND_BANDS",
                     "EDISP", "EDISP_BANDS", "PSF", "PSF_BANDS", "MASK_SAFE", "MASK_SAFE_BANDS", "MASK_FIT", "MASK_FIT_BANDS"]
    back = MapDataset.read(path)
    assert back.name == "obs-7"
    for attr in ("counts", "exposure", "background"):
        _assert_same_map(getattr(ds, attr), getattr(back, attr))
    assert back.exposure.geom.axes[0].name == "energy_true" and back.counts.geom.axes[0].name == "energy"
    for attr in ("mask_safe", "mask_fit"):
        assert getattr(back, attr).data.dtype == bool
        _assert_same_map(getattr(ds, attr), getattr(back, attr))
    assert back.psf.has_single_spatial_bin == (not zoned) and back.edisp.has_single_spatial_bin == (not zoned)
    np.testing.assert_array_equal(back.psf.data, ds.psf.data)
    np.testing.assert_array_equal(back.psf.rad_axis.edges, ds.psf.rad_axis.edges)
    if zoned:
        assert back.psf.geom == ds.psf.geom and back.edisp.geom == ds.edisp.geom
        np.testing.assert_array_equal(back.edisp.data, ds.edisp.data)
    else:
        np.testing.assert_array_equal(back.edisp.get_edisp_kernel().pdf_matrix, ds.edisp.get_edisp_kernel().pdf_matrix)
    # bands table of the exposure: the reference's column names for a true-energy axis
    hdus = {h.name: h for h in fitsio.read_hdus(path)}
    assert set(hdus["EXPOSURE_BANDS"].data) == {"CHANNEL", "ENERGY_TRUE", "ENERGY_TRUE_MIN", "ENERGY_TRUE_MAX"}
    assert set(hdus["COUNTS_BANDS"].data) == {"CHANNEL", "ENERGY", "E_MIN", "E_MAX"}
    assert hdus["COUNTS"].header["BANDSHDU"] == "COUNTS_BANDS" and hdus["COUNTS"].header["CTYPE1"] == "GLON-CAR"
    assert hdus["PSF_BANDS"].data["CHANNEL"].size == ds.psf.data.shape[0] * ds.psf.data.shape[1]


@pytest.mark.parametrize("proj,skydir", [("CAR", (83.63, 22.01)), ("CAR", (266.4, -29.0)), ("TAN", (30.0, 40.0))])
def test_off_equator_geometry_round_trip(tmp_path, proj, skydir):
    """CRVAL2, the projection code and the pole keywords of an oblique CAR / TAN image survive the round trip; the pole
    keywords are the defaults of FITS WCS paper II (the only ones the kernels implement), anything else is refused."""
    ds = configs.zoned_irfs(configs.make_dataset(width=(3.0, 1.6), n_reco=5, n_true=8, mask_radius=0.7, skydir=skydir,
                                                 proj=proj, name="off"))
    path = str(tmp_path / "off.fits")
    ds.write(path)
    back = MapDataset.read(path)
    assert back._geom == ds._geom and back._geom.projection == proj and back._geom.crval == ds._geom.crval
    assert back.psf.geom == ds.psf.geom and back.psf.geom.projection == proj
    _assert_same_map(ds.exposure, back.exposure)
    a, b = ds._geom.to_image().get_coord(), back._geom.to_image().get_coord()
    np.testing.assert_array_equal(a["lon"], b["lon"])
    np.testing.assert_array_equal(a["lat"], b["lat"])
    hdus = fitsio.read_hdus(path)
    h = {x.name: x for x in hdus}["COUNTS"].header
    assert h["CTYPE1"] == ("GLON-" + proj) and h["CRVAL2"] == ds._geom.crval[1]
    if proj == "CAR":
        assert h["LONPOLE"] == (0.0 if skydir[1] >= 0 else 180.0) and h["LATPOLE"] == pytest.approx(90.0 - abs(skydir[1]))
    else:
        assert h["LONPOLE"] == 180.0 and h["LATPOLE"] == pytest.approx(skydir[1])
    # a rotated image (non-default LONPOLE) is not something the kernels can evaluate: refuse to read it
    for x in hdus:
        if x.name == "COUNTS":
            x.header["LONPOLE"] = 90.0
    fitsio.write_hdus(str(tmp_path / "rot.fits"), hdus)
    with pytest.raises(NotImplementedError):
        MapDataset.read(str(tmp_path / "rot.fits"))
