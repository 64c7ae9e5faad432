"""CPU oracle for the TTDG-MGM test-time-adaptation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and there only as the checker or the
timed CPU baseline - never as the thing shipped.  The product path
(``ttdg-mgm_b200/``) never imports this package and fails loudly when its CUDA
library is missing.

Parity status (see DESIGN.md "Oracle"):

* MGM stage (``mgm_port.py``)      - pinned against the reference's own Python
  (``/root/reference/adapteacher/modeling/GModule``) imported unmodified through
  ``ref_shim.py``; golden vectors in ``tests/golden/`` were produced by
  ``gen_golden.py`` from that import.
* ``pygmtools.sinkhorn`` (``pygm_sinkhorn.py``) - third-party, pinned
  ``pygmtools==0.3.8`` (reference ``requirements.txt:58``), NOT present in this
  image: restated from its published algorithm -> **parity unpinned**.
* ``scipy.optimize.linear_sum_assignment`` (``lap_ref.c``) - third-party; SciPy
  1.18.1 IS present and the C restatement is checked against it (reference pins
  1.7.3) -> pinned against the installed SciPy only.
* Detectron2 0.5 detector (``detector_port.py``) - third-party, not installable
  here: restated from SURVEY.md Appendix A -> **parity unpinned**.
"""
