"""CPU oracle for the gammapy forward-fold likelihood hot path.

TEST INFRASTRUCTURE ONLY.  This package is a numpy/scipy restatement of the
reference algorithm (``MapEvaluator.compute_npred`` + Cash) used as the
checker in ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py``.  Nothing under ``gammapy_b200/``
imports it; the product path runs on the CUDA library only.

Every function cites the reference ``file:line`` (relative to
``/root/reference/gammapy``) it follows.  Parity pinning:

* Cash sum: pinned against the reference's own Cython kernel compiled
  unmodified into ``oracle/_ref`` (``oracle/build_ref.py``) and the Sherpa
  derived per-bin goldens of ``stats/tests/test_fit_statistics.py``.
* Spectral integrals: pinned by ``modeling/models/tests/test_spectral.py``
  ``TEST_MODELS`` and ``utils/tests/test_integrate.py``.
* Spatial models / WCS / solid angle / oversampling: pinned by
  ``test_evaluate_integrate_nd_geom`` (``models/tests/test_cube.py:802-845``),
  ``test_sky_disk`` / ``test_sky_gaussian`` (``models/tests/test_spatial.py``)
  and the ``Cutout2D`` index rule by ``maps/wcs/tests/test_geom.py:279-329``.
* Full forward fold: pinned by ``test_npred_psf_after_edisp``
  (``datasets/tests/test_map.py:1418-1445``, sum = 129553.858658).
* Fit-level numbers (``test_map_fit``) need ``$GAMMAPY_DATA`` + iminuit and are
  therefore **parity unpinned**; static IRF builders (``get_psf_kernel`` /
  ``EDispKernel.from_gauss``) are pinned only through the npred total above.

The python reference itself cannot be imported here (astropy / regions /
iminuit are not installed), so astropy/wcslib arithmetic (Vincenty separation,
position angle, CAR projection, ``Cutout2D`` slices) is restated from the
published formulas; see ``oracle/wcs.py``.
"""
