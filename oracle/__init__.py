"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the embodied_b200 hot path.

Nothing under ``oracle/`` is product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker (or as the
timed CPU baseline), never as the thing shipped.

Contents
  _shim/elements.py, _shim/portal.py
      From-scratch mini re-implementations of the two third-party packages the
      reference's ``embodied/core`` imports (``elements>=3.17``, ``portal>=3.5``,
      ``requirements.txt:5,15``; neither is vendored under /root/reference).
      Only the bookkeeping surface listed in SURVEY.md section 8c.
  refload.py
      Loads the reference's OWN ``embodied/core/*.py`` verbatim, by path, from
      /root/reference (this container only; the GPU box has no /root/reference).
      Used to pin ``host_oracle`` and to generate ``tests/golden``.
  host_oracle.py
      numpy restatement of Driver / Replay / Chunk / Uniform / Consec, each
      function citing the reference file:line it follows.  Travels to the GPU
      box.  PARITY PINNED: checked against the real reference (refload) in
      ``tests/test_oracle_pinned.py`` and against ``tests/golden/*.npz``.
  dreamer_oracle.py
      fp32 torch-CPU restatement of the dreamerv3 math on the path.
      PARITY UNPINNED: the JAX original cannot run here (jax/ninjax/optax are not
      installed) and the reference ships no golden tensors for it.
  gen_golden.py
      The script that produced ``tests/golden`` from the real reference.
"""
