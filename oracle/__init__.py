"""TEST INFRASTRUCTURE ONLY.

CPU oracle for the caustics hot path.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this package; nothing under
``caustics_b200/`` (the product) imports, links or executes it.

  oracle.solver   -- ctypes front ends: ``ref_solve`` (the UNMODIFIED reference custom call built
                     from /root/reference into oracle/_ref/) and ``port_solve`` (ea_oracle.c, our
                     plain-C restatement).
  oracle.lens     -- NumPy restatement of the point-source layer (coefficients, lens equation,
                     image filter, Jacobian magnification, hexadecapole, gate).
  oracle.extended -- NumPy restatement of the contour-integration extended-source pipeline.
  oracle.refshim  -- (this container only) runs the reference's own Python under a NumPy stand-in
                     for jax to generate the golden vectors in tests/golden/.
"""
