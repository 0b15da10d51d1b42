"""CPU oracle for the FABOLAS GP inner loop.

TEST INFRASTRUCTURE ONLY.  Nothing under ``fabolas_b200/`` may import this package; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs do, and there only as the checker (or as the timed CPU baseline), never as the product.

Pinning status (see DESIGN.md §"Oracle"):
  * ``epmgp_oracle`` / ``acq_oracle`` / ``priors_oracle`` / ``ard_oracle`` are PINNED against the
    reference's own ``epmgp.py`` / ``acquisitions.py`` / ``horseshoe.py`` / ``ard.py`` run in the
    authoring container; the outputs are committed under ``tests/golden/`` together with the
    generating script ``tests/golden/make_golden.py``.
  * ``george_oracle`` restates george 0.4.0 (third party, pinned in the reference's
    ``requirements.txt:6`` but absent from ``/root/reference`` and not installable offline).
    PARITY UNPINNED against george itself; it is cross-checked against scikit-learn's independent
    GaussianProcessRegressor (Matern/amplitude/noise part) and central finite differences.
"""
