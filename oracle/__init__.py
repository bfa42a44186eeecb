"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the LightHand heatmap pipeline (crop/warp+normalise ->
Gaussian targets -> JointsMSELoss -> argmax decode -> PCK/AUC/EPE), used as the
checker for the CUDA path in ``lighthand_b200``.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this package.  The product
package ``lighthand_b200`` never imports it and has no CPU fallback.

Parity pinning: the reference ships no golden vectors.  The
restatement in ``oracle/restate.py`` is pinned against the reference's *own
functions executed in the build container* (``oracle/reference_bridge.py`` +
``tests/golden/make_golden.py`` -> committed fixtures under ``tests/golden``).
Sub-steps that the reference does not implement at all (quarter-pixel shift,
inverse affine inside decode, target_weight-masked loss, geometry other than
64x64 / sigma 2 / stride 4) are restated from upstream HRNet semantics and are
"parity unpinned"; each such function says so in its docstring.
"""
