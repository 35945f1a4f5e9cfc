"""CPU oracle for the CFG-DDPM sampling path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it, and there only as the checker / CPU baseline.
The product path (``spectrogramgenai_b200``) never imports this package and
fails loudly when its CUDA library is missing.

Parity status: the reference's own tests do not pin this path
(``/root/reference/tests/test_main.py`` covers ``helpers.fast_resize_m1_1``
only).  The oracle is therefore pinned against OUTPUTS OF THE REFERENCE ITSELF,
run in the build container by ``tests/golden/make_golden.py`` (script committed,
vectors committed under ``tests/golden/``); ``tests/test_oracle_golden.py``
re-checks the oracle against those vectors on every run.
"""
