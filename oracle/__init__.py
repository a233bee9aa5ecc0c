"""CPU oracle for the MUFASA per-pixel spectral-fitting hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``mufasa_b200/`` imports this package; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may use it, and only as the checker / timed CPU baseline -- never as the product path.

It is a numpy restatement of what the reference computes on the path
``UltraCube.fit_cube -> multi_v_fit.cubefit_gen/cubefit_simp -> pyspeckit.Cube.fiteach -> mpfit``
plus the statistics half (``get_rss / expand_mask / AICc / likelihood / get_best_2c_parcube``).

PARITY STATUS (see DESIGN.md "Oracle"):
  * model, AIC/AICc/likelihood, get_rss/expand_mask/get_rms/get_chisq, moment_guesses,
    tautex_renorm: PINNED -- checked against golden vectors produced by executing the
    reference's own source files (tests/golden/make_golden.py imports them from /root/reference
    with stub modules standing in for the absent astropy/pyspeckit/spectral_cube imports).
  * LM solver (pyspeckit ``Cube.fiteach`` + ``mpfit``), the NH3(1,1) hyperfine table:
    PARITY UNPINNED -- third-party, un-vendored (``pyspeckit>=1.0.4``, pyproject.toml:22), cannot be
    installed here; restated from the published MINPACK-1 lmdif algorithm + Markwardt's
    box-limit extension and cross-checked against scipy's real MINPACK (``scipy.optimize.leastsq``).
"""
