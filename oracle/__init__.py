"""oracle/ -- CPU restatements of the reference's hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import, call, link or execute anything in this package, and there only as the checker.  The
product package (taxidispatcher_b200/) never imports it and fails loudly when its CUDA library
is missing.

Modules
  gen_inputs  deterministic synthetic inputs for every BASELINE.json config
  cost_ref    calculate_cost restated (split.py:123-136, simulate.py:17-33, procedure.py:6-12)
  lcm_ref     the reference LCM bodies lifted verbatim-in-behaviour (numpy) + a C twin
  assign_ref  exact optimum (C shortest-augmenting-path), scipy LSA and LP-relaxation cross-checks
  pool_ref    pool finder: C restatement + driver for the compiled reference (oracle/_ref/pool_n_big)
  _clib       builds/loads oracle/_build/liboracle.so
"""
