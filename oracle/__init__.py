"""CPU oracle for the AE + RaPP hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this package.  The
product path (``icra2021_multimodal_ad_b200``) never imports it and has no CPU
fallback.

Parity status: PINNED.  Every function is checked (``tests/test_oracle_golden.py``)
against fixtures under ``tests/golden/`` that were produced by importing and
running the unmodified reference (``/root/reference``) in the build container
with ``tests/golden/make_golden.py`` (committed).  The reference ships no golden
vectors or tests of its own, so those generated outputs are
the pin.  sklearn 1.9.0 / numpy 2.3.5 / torch 2.11.0 are the third-party
libraries whose arithmetic the metric restatements follow.
"""
from . import rapp_oracle, metric_oracle  # noqa: F401
