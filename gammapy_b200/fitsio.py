"""Dataset ingestion: a minimal FITS reader / writer and the GADF map-dataset layout on top of it.

astropy is not available here, so this is a from-scratch implementation of the part of the FITS standard the layout
needs: a primary HDU, image extensions (BITPIX 8 / 16 / 32 / 64 / -32 / -64, big endian) and binary tables with scalar
columns (L, B, I, J, K, E, D and nA strings).  The layout follows what the reference writes:

* ``MapDataset.to_hdulist`` / ``from_hdulist``   datasets/map.py:1556-1700: HDUs COUNTS, EXPOSURE, BACKGROUND, EDISP, PSF,
  MASK_SAFE, MASK_FIT, each followed by its ``<NAME>_BANDS`` table; NAME in the primary header
* ``WcsNDMap.to_hdu`` / ``Map.from_hdulist``     maps/wcs/ndmap.py, maps/core.py: image in (..., lat, lon) order, ``BANDSHDU``
  pointing at the bands table
* ``WcsGeom.to_header``                          maps/wcs/geom.py:544-552: WCS keywords of the image plane + ``WCSSHAPE``
* ``MapAxes.to_table`` / ``to_header``           maps/axes.py:1963-2080: CHANNEL + <AXIS> / <AXIS>_MIN / <AXIS>_MAX columns
  (ENERGY / E_MIN / E_MAX for the axis called "energy"), ``AXCOLSn`` / ``INTERPn`` keywords

Files written here and read back are covered by tests (round trip, byte-level structure); the reference's own reader
could not be run against them (no astropy in this environment): interoperability is by construction, unverified.
"""
import numpy as np

BLOCK = 2880

_BITPIX = {np.dtype("uint8"): 8, np.dtype(">i2"): 16, np.dtype(">i4"): 32, np.dtype(">i8"): 64,
           np.dtype(">f4"): -32, np.dtype(">f8"): -64}
_FROM_BITPIX = {8: "uint8", 16: ">i2", 32: ">i4", 64: ">i8", -32: ">f4", -64: ">f8"}
_TFORM = {"L": ("S1", 1), "B": ("u1", 1), "I": (">i2", 2), "J": (">i4", 4), "K": (">i8", 8), "E": (">f4", 4), "D": (">f8", 8)}


class HDU:
    """One header-data unit: ``data`` is a numpy array (image) or a dict name -> 1-D array (binary table) or None."""

    def __init__(self, name, data=None, header=None, units=None):
        self.name = name.upper()
        self.data = data
        self.header = dict(header or {})
        self.units = dict(units or {})  # table columns only: name -> TUNIT

    @property
    def is_table(self):
        return isinstance(self.data, dict)


# ---- writing ---------------------------------------------------------------------------------------------------------
def _card(key, value, comment=""):
    if isinstance(value, bool):
        v = f"{'T' if value else 'F':>20}"
    elif isinstance(value, (int, np.integer)):
        v = f"{int(value):>20d}"
    elif isinstance(value, (float, np.floating)):
        text = repr(float(value)).upper()
        if "E" not in text and "." not in text and "N" not in text:
            text += "."
        v = f"{text:>20}"
    else:
        text = str(value).replace("'", "''")
        v = f"'{text:<8}'"
        if len(v) > 70:
            raise ValueError(f"FITS string value of {key} too long")
    card = f"{key:<8}= {v}"
    if comment:
        card += f" / {comment}"
    if len(card) > 80:
        card = card[:80]
    return f"{card:<80}"


def _pad(raw, fill):
    n = (-len(raw)) % BLOCK
    return raw + fill * n


def _header_bytes(cards):
    text = "".join(cards) + f"{'END':<80}"
    return _pad(text.encode("ascii"), b" ")


def _image_bytes(hdu, primary):
    data = hdu.data
    cards = []
    if primary:
        cards.append(_card("SIMPLE", True, "conforms to FITS standard"))
    else:
        cards.append(_card("XTENSION", "IMAGE", "Image extension"))
    if data is None:
        cards += [_card("BITPIX", 8), _card("NAXIS", 0)]
        payload = b""
    else:
        arr = np.asarray(data)
        if arr.dtype == bool:
            arr = arr.astype("uint8")  # the reference writes boolean masks as uint8 images
        be = arr.dtype.newbyteorder(">") if arr.dtype.itemsize > 1 else arr.dtype
        if be not in _BITPIX:
            raise ValueError(f"unsupported image dtype {arr.dtype}")
        arr = np.ascontiguousarray(arr, dtype=be)
        cards += [_card("BITPIX", _BITPIX[be]), _card("NAXIS", arr.ndim)]
        for i, n in enumerate(arr.shape[::-1]):
            cards.append(_card(f"NAXIS{i + 1}", n))
        payload = _pad(arr.tobytes(), b"\0")
    if primary:
        cards.append(_card("EXTEND", True))
    else:
        cards += [_card("PCOUNT", 0), _card("GCOUNT", 1), _card("EXTNAME", hdu.name)]
    for k, v in hdu.header.items():
        cards.append(_card(k, v))
    return _header_bytes(cards) + payload


def _table_bytes(hdu):
    cols = {k: np.asarray(v) for k, v in hdu.data.items()}
    nrow = len(next(iter(cols.values()))) if cols else 0
    fields, forms = [], []
    for name, arr in cols.items():
        if arr.ndim != 1 or len(arr) != nrow:
            raise ValueError("binary-table columns must be 1-D arrays of one length")
        if arr.dtype.kind in "US":
            width = max(int(arr.dtype.itemsize // (4 if arr.dtype.kind == "U" else 1)), 1)
            fields.append((name, f"S{width}"))
            forms.append(f"{width}A")
        elif arr.dtype.kind == "b":
            fields.append((name, "S1"))
            forms.append("L")
        elif arr.dtype.kind == "f":
            code = "E" if arr.dtype.itemsize == 4 else "D"
            fields.append((name, _TFORM[code][0]))
            forms.append(code)
        elif arr.dtype.kind in "iu":
            code = {1: "B", 2: "I", 4: "J", 8: "K"}[arr.dtype.itemsize]
            fields.append((name, _TFORM[code][0]))
            forms.append(code)
        else:
            raise ValueError(f"unsupported column dtype {arr.dtype}")
    rec = np.zeros(nrow, dtype=np.dtype(fields))
    for (name, _), form in zip(fields, forms):
        arr = cols[name]
        if form == "L":
            rec[name] = np.where(arr, b"T", b"F")
        elif form.endswith("A"):
            rec[name] = np.char.encode(arr.astype(str), "ascii")
        else:
            rec[name] = arr
    cards = [_card("XTENSION", "BINTABLE", "binary table extension"), _card("BITPIX", 8), _card("NAXIS", 2),
             _card("NAXIS1", rec.dtype.itemsize), _card("NAXIS2", nrow), _card("PCOUNT", 0), _card("GCOUNT", 1),
             _card("TFIELDS", len(fields)), _card("EXTNAME", hdu.name)]
    for i, ((name, _), form) in enumerate(zip(fields, forms)):
        cards += [_card(f"TTYPE{i + 1}", name), _card(f"TFORM{i + 1}", form)]
        if hdu.units.get(name):
            cards.append(_card(f"TUNIT{i + 1}", hdu.units[name]))
    for k, v in hdu.header.items():
        cards.append(_card(k, v))
    return _header_bytes(cards) + _pad(rec.tobytes(), b"\0")


def write_hdus(path, hdus, overwrite=False):
    """Write ``hdus`` (the first one is the primary HDU) as one FITS file."""
    import os

    if os.path.exists(path) and not overwrite:
        raise OSError(f"{path} exists; use overwrite=True")
    with open(path, "wb") as fh:
        for i, hdu in enumerate(hdus):
            fh.write(_table_bytes(hdu) if hdu.is_table else _image_bytes(hdu, primary=i == 0))


# ---- reading ---------------------------------------------------------------------------------------------------------
def _parse_value(text):
    text = text.strip()
    if text.startswith("'"):
        end = 1
        out = []
        while end < len(text):
            if text[end] == "'":
                if end + 1 < len(text) and text[end + 1] == "'":
                    out.append("'")
                    end += 2
                    continue
                break
            out.append(text[end])
            end += 1
        return "".join(out).rstrip()
    text = text.split("/")[0].strip()
    if text == "T":
        return True
    if text == "F":
        return False
    try:
        return int(text)
    except ValueError:
        return float(text.replace("D", "E"))


def _read_header(fh):
    cards = {}
    order = []
    while True:
        block = fh.read(BLOCK)
        if len(block) < BLOCK:
            return None, None
        done = False
        for i in range(0, BLOCK, 80):
            card = block[i:i + 80].decode("ascii")
            key = card[:8].strip()
            if key == "END":
                done = True
                break
            if card[8:10] == "= ":
                cards[key] = _parse_value(card[10:])
                order.append(key)
        if done:
            return cards, order


def read_hdus(path):
    """All HDUs of a FITS file as a list of ``HDU`` (images in numpy (..., y, x) order, tables as column dicts)."""
    out = []
    structural = ("SIMPLE", "XTENSION", "BITPIX", "NAXIS", "PCOUNT", "GCOUNT", "EXTEND", "EXTNAME", "TFIELDS")
    with open(path, "rb") as fh:
        while True:
            cards, order = _read_header(fh)
            if cards is None:
                break
            naxis = cards.get("NAXIS", 0)
            shape = [cards[f"NAXIS{i + 1}"] for i in range(naxis)]
            nbytes = abs(cards["BITPIX"]) // 8 * int(np.prod(shape)) if naxis else 0
            raw = fh.read(nbytes + (-nbytes) % BLOCK)[:nbytes]
            name = cards.get("EXTNAME", "PRIMARY")
            header = {k: cards[k] for k in order if k not in structural and not k.startswith(("NAXIS", "TTYPE", "TFORM", "TUNIT"))}
            if cards.get("XTENSION", "").strip() == "BINTABLE":
                fields, kinds, units = [], [], {}
                for i in range(cards["TFIELDS"]):
                    form = cards[f"TFORM{i + 1}"].strip()
                    col = cards[f"TTYPE{i + 1}"].strip()
                    if form.endswith("A"):
                        fields.append((col, f"S{int(form[:-1] or 1)}"))
                    else:
                        code = form.lstrip("1")
                        if code not in _TFORM:
                            raise ValueError(f"unsupported TFORM {form!r} in HDU {name}")
                        fields.append((col, _TFORM[code][0]))
                    kinds.append(form)
                    if f"TUNIT{i + 1}" in cards:
                        units[col] = cards[f"TUNIT{i + 1}"]
                rec = np.frombuffer(raw, dtype=np.dtype(fields), count=cards["NAXIS2"])
                data = {}
                for (col, _), form in zip(fields, kinds):
                    if form == "L":
                        data[col] = rec[col] == b"T"
                    elif form.endswith("A"):
                        data[col] = np.char.decode(rec[col], "ascii")
                    else:
                        data[col] = rec[col].astype(rec[col].dtype.newbyteorder("="))
                out.append(HDU(name, data, header, units))
            elif naxis:
                dt = np.dtype(_FROM_BITPIX[cards["BITPIX"]])
                arr = np.frombuffer(raw, dtype=dt).reshape(shape[::-1])
                arr = arr.astype(dt.newbyteorder("=")) * cards.get("BSCALE", 1) + cards.get("BZERO", 0) \
                    if ("BSCALE" in cards or "BZERO" in cards) else arr.astype(dt.newbyteorder("="))
                out.append(HDU(name, arr, header))
            else:
                out.append(HDU(name, None, header))
    return out


# ---- maps <-> HDUs ---------------------------------------------------------------------------------------------------------
_CTYPE = {"galactic": ("GLON-CAR", "GLAT-CAR"), "icrs": ("RA---CAR", "DEC--CAR")}
_CTYPE_TAN = {"galactic": ("GLON-TAN", "GLAT-TAN"), "icrs": ("RA---TAN", "DEC--TAN")}


def _axis_columns(ax):
    name = ax.name.upper()
    return ("ENERGY", "E_MIN", "E_MAX") if name == "ENERGY" else (name, name + "_MIN", name + "_MAX")


def _default_lonpole(proj, crval_lat):
    """LONPOLE when the header does not give it (FITS WCS paper II section 2.4): 180 deg for a zenithal projection
    (theta_0 = 90 deg), for CAR (theta_0 = 0) 0 deg if delta_0 >= theta_0 else 180 deg.  The only value the kernels know."""
    if proj == "TAN":
        return 180.0
    return 0.0 if crval_lat >= 0.0 else 180.0


def map_to_hdus(m, hdu):
    """``WcsNDMap.to_hdulist(hdu=...)`` without the primary: the image HDU and, for a cube, its BANDS table."""
    geom = m.geom
    hdu = hdu.upper()
    ctype = (_CTYPE_TAN if geom.projection == "TAN" else _CTYPE).get(geom.frame)
    if ctype is None:
        raise NotImplementedError(f"FITS header for frame {geom.frame!r}")
    header = {"WCSAXES": 2, "CRPIX1": geom.crpix[0], "CRPIX2": geom.crpix[1], "CDELT1": -geom.binsz, "CDELT2": geom.binsz,
              "CUNIT1": "deg", "CUNIT2": "deg", "CTYPE1": ctype[0], "CTYPE2": ctype[1], "CRVAL1": geom.crval[0],
              "CRVAL2": geom.crval[1], "LONPOLE": _default_lonpole(geom.projection, geom.crval[1]),
              # the native pole's latitude: the reference point for TAN, 90 - |delta_0| on the reference meridian for CAR
              "LATPOLE": geom.crval[1] if geom.projection == "TAN" else 90.0 - abs(geom.crval[1])}
    shape = ",".join(str(n) for n in (geom.npix[0], geom.npix[1]) + tuple(ax.nbin for ax in geom.axes))
    header["WCSSHAPE"] = f"({shape})"
    if getattr(m, "unit", ""):
        header["BUNIT"] = str(m.unit)
    out = []
    if len(geom.axes):
        bands = f"{hdu}_BANDS"
        header["BANDSHDU"] = bands
        grids = np.meshgrid(*[np.arange(ax.nbin) for ax in geom.axes])  # as the reference: np.meshgrid of the axes
        cols = {"CHANNEL": np.arange(int(np.prod([ax.nbin for ax in geom.axes])), dtype=np.int64)}
        units = {}
        for i, ax in enumerate(geom.axes):
            idx = np.ravel(grids[i])
            c_ctr, c_min, c_max = _axis_columns(ax)
            cols[c_ctr], cols[c_min], cols[c_max] = ax.center[idx], ax.edges_min[idx], ax.edges_max[idx]
            for c in (c_ctr, c_min, c_max):
                units[c] = ax.unit
            header[f"AXCOLS{i + 1}"] = f"{c_min},{c_max}"
            header[f"INTERP{i + 1}"] = ax.interp
        out.append(HDU(bands, cols, {k: v for k, v in header.items() if k.startswith(("AXCOLS", "INTERP"))}, units))
    return [HDU(hdu, np.asarray(m.data), header)] + out


def map_from_hdus(hdus, hdu, axis_names=None):
    """``Map.from_hdulist(hdulist, hdu=...)``: geometry from the WCS keywords and the BANDS table."""
    from .maps import Map, MapAxis, WcsGeom

    by_name = {h.name: h for h in hdus}
    image = by_name[hdu.upper()]
    h = image.header
    ctype1 = str(h.get("CTYPE1", "")).strip()
    proj = "TAN" if ctype1.endswith("TAN") else "CAR"
    frame = {v[0]: k for k, v in (_CTYPE_TAN if proj == "TAN" else _CTYPE).items()}.get(ctype1)
    if frame is None:
        raise NotImplementedError(f"projection / frame {h.get('CTYPE1')!r} (CAR or TAN images in galactic or icrs only)")
    if abs(abs(h["CDELT1"]) - abs(h["CDELT2"])) > 1e-12 * abs(h["CDELT2"]):
        raise NotImplementedError("square pixels only")
    if "LONPOLE" in h and abs((float(h["LONPOLE"]) - _default_lonpole(proj, float(h["CRVAL2"])) + 180.0) % 360.0 - 180.0) > 1e-9:
        raise NotImplementedError(f"LONPOLE = {h['LONPOLE']} is not the default of this projection (a rotated image)")
    data = image.data
    ny, nx = data.shape[-2:]
    axes = []
    if "BANDSHDU" in h:
        table = by_name[str(h["BANDSHDU"]).strip().upper()]
        n_axes = data.ndim - 2
        remaining = int(np.prod(data.shape[:-2]))
        stride = 1
        for i in range(n_axes):  # axis i varies fastest-first in the raveled meshgrid of the writer
            cols = str(table.header.get(f"AXCOLS{i + 1}", h.get(f"AXCOLS{i + 1}", "E_MIN,E_MAX"))).split(",")
            lo, hi = table.data[cols[0].strip()], table.data[cols[1].strip()]
            nbin = data.shape[-3 - i]
            if n_axes == 1:
                sel = np.arange(nbin)
            else:  # np.meshgrid(ax0, ax1) raveled: index = i1 * n0 + i0 for two axes ("xy" indexing)
                sel = np.arange(nbin) * stride if i == 0 else np.arange(nbin) * stride
            edges = np.append(lo[sel], hi[sel][-1])
            name = cols[0].strip()[:-4].lower()
            name = "energy" if name == "e" else name
            interp = str(table.header.get(f"INTERP{i + 1}", h.get(f"INTERP{i + 1}", "log" if name.startswith("energy") else "lin")))
            unit = table.units.get(cols[0].strip(), "")
            axes.append(MapAxis(edges, interp=interp.strip(), name=name, node_type="edges", unit=unit))
            stride *= nbin
        del remaining
    if axis_names:
        for ax, new in zip(axes, axis_names):
            ax.name = new
    geom = WcsGeom((nx, ny), abs(h["CDELT2"]), (h["CRVAL1"], h["CRVAL2"]), (h["CRPIX1"], h["CRPIX2"]), frame=frame,
                   axes=axes, proj=proj)
    m = Map.from_geom(geom, data=data, dtype=data.dtype, unit=str(h.get("BUNIT", "")))
    return m


# ---- MapDataset <-> file -----------------------------------------------------------------------------------------------------
def _single_pixel_geom(geom):
    """Image geometry of one pixel covering the dataset's field: the spatial grid of a position-independent IRF map."""
    from .maps import WcsGeom

    c = geom.center_skydir
    return WcsGeom.create(skydir=(c.lon, 0.0), binsz=float(max(geom.width)), npix=(1, 1), frame=geom.frame, proj="CAR")


def dataset_to_hdus(ds):
    """``MapDataset.to_hdulist`` (datasets/map.py:1556-1607); models are not serialised (neither does the reference)."""
    from .maps import Map

    hdus = [HDU("PRIMARY", None, {"NAME": ds.name})]
    for attr in ("counts", "exposure", "background"):
        m = getattr(ds, attr)
        if m is not None:
            hdus += map_to_hdus(m, attr)
    for attr in ("counts_off", "acceptance", "acceptance_off"):  # MapDatasetOnOff.to_hdulist (datasets/map.py:3038-3075)
        m = getattr(ds, attr, None)
        if m is not None:
            hdus += map_to_hdus(m, attr)
    geom = ds._geom.to_image()
    if ds.edisp is not None:
        ed = ds.edisp
        if ed.has_single_spatial_bin:
            axes = [ed.kernel.axes["energy"], ed.kernel.axes["energy_true"]]
            g = _single_pixel_geom(geom).to_cube(axes)
            data = np.asarray(ed.kernel.pdf_matrix, dtype=float)[:, :, None, None]
        else:
            g, data = ed.geom.to_cube([ed.axes[1], ed.axes[0]]), ed.data
        hdus += map_to_hdus(Map.from_geom(g, data=data, dtype=float), "edisp")
    if ds.psf is not None and hasattr(ds.psf, "rad_axis"):
        psf = ds.psf
        g0 = _single_pixel_geom(geom) if psf.has_single_spatial_bin else psf.geom
        data = psf.data[:, :, None, None] if psf.has_single_spatial_bin else psf.data
        hdus += map_to_hdus(Map.from_geom(g0.to_cube([psf.rad_axis, psf.energy_axis_true]), data=data, dtype=float,
                                          unit="sr-1"), "psf")
    for attr in ("mask_safe", "mask_fit"):
        m = getattr(ds, attr)
        if m is not None:
            hdus += map_to_hdus(Map.from_geom(m.geom, data=np.asarray(m.data).astype(np.uint8), dtype=np.uint8), attr)
    return hdus


def dataset_from_hdus(hdus, name=None):
    """``MapDataset.from_hdulist`` (datasets/map.py:1609-1700)."""
    from .datasets import MapDataset
    from .irf import EDispKernel, EDispKernelMap, PSFMap

    names = {h.name for h in hdus}
    kwargs = {"name": name if name is not None else hdus[0].header.get("NAME")}
    if "COUNTS" in names:
        kwargs["counts"] = map_from_hdus(hdus, "counts")
    if "EXPOSURE" in names:
        exposure = map_from_hdus(hdus, "exposure")
        if exposure.geom.axes[0].name == "energy":  # as the reference: an exposure cube is in true energy
            exposure.geom.axes[0].name = "energy_true"
        kwargs["exposure"] = exposure
    if "BACKGROUND" in names:
        kwargs["background"] = map_from_hdus(hdus, "background")
    if "EDISP" in names:
        m = map_from_hdus(hdus, "edisp")
        e_reco, e_true = m.geom.axes[0], m.geom.axes[1]
        if m.data.shape[-2:] == (1, 1):
            kwargs["edisp"] = EDispKernelMap(EDispKernel([e_true, e_reco], m.data[:, :, 0, 0]))
        else:
            kwargs["edisp"] = EDispKernelMap(data=np.asarray(m.data, float), geom=m.geom.to_image(), axes=[e_true, e_reco])
    if "PSF" in names:
        m = map_from_hdus(hdus, "psf")
        rad, e_true = m.geom.axes[0], m.geom.axes[1]
        if m.data.shape[-2:] == (1, 1):
            kwargs["psf"] = PSFMap(e_true, rad, np.asarray(m.data, float)[:, :, 0, 0])
        else:
            kwargs["psf"] = PSFMap(e_true, rad, np.asarray(m.data, float), geom=m.geom.to_image())
    for attr in ("mask_safe", "mask_fit"):
        if attr.upper() in names:
            m = map_from_hdus(hdus, attr)
            m.data = m.data.astype(bool)
            kwargs[attr] = m
    if "COUNTS_OFF" in names or "ACCEPTANCE" in names:  # MapDatasetOnOff.from_hdulist (datasets/map.py:3077-3160)
        from .datasets import MapDatasetOnOff

        for attr in ("counts_off", "acceptance", "acceptance_off"):
            if attr.upper() in names:
                kwargs[attr] = map_from_hdus(hdus, attr)
        kwargs.pop("background", None)
        return MapDatasetOnOff(**kwargs)
    return MapDataset(**kwargs)
