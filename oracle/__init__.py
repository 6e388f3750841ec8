"""CPU oracle for the DCARL hot path — TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  ``dcarl_b200`` never does.

Two layers, both pinned against outputs of the reference itself (tests/golden/,
made by tools/make_golden.py which runs the unmodified reference scripts):

* ``oracle.pyref``  — a literal Python/NumPy restatement of the reference loops
  (same list-append / np.array(list) / np.mean / np.std calls), slow like the
  reference, bit-exact by construction.
* ``oracle.cport``  — ctypes bindings to ``dcarl_oracle.c`` (plain C restatement
  including numpy's pairwise summation), bit-exact with ``pyref`` and ~100x faster,
  used for large parity samples and the CPU timing legs.
"""
