"""CPU oracle for the RSTnet real-time inference hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``rstnet_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline /
``--impl reference`` legs of ``bench.py`` use it, and there only as the
checker / the reported CPU baseline, never as the product path.

The oracle is a functional restatement (torch CPU fp32 ops, the same ATen
calls the pure-Python reference makes) of the reference algorithms, each
function citing the reference file:line it follows.  It is pinned against the
unmodified reference run in the build container by ``oracle/gen_golden.py``
(fixtures under ``tests/golden/``).
"""
