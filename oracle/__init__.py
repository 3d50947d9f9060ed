"""CPU oracle for the PPO + TransformerXL hot path.  TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (torch-CPU fp32 ops, functional style, parameters passed as a
plain ``{state_dict_name: tensor}`` dict) of the algorithm that
MarcoMeter/episodic-transformer-memory-ppo implements in ``transformer.py``, ``model.py``, ``buffer.py``,
``utils.py`` and ``trainer.py``.  Every function cites the reference file:line it follows.

Pinning: the reference has no tests and no golden vectors of its own.  The oracle
is pinned by executing the *unmodified* reference modules in the build container
(``tests/golden/make_golden.py`` imports them from ``/root/reference`` behind import stubs) and
committing their inputs/outputs as fixtures under ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks this oracle against every one of them.

Who may import this package: ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` -- as the checker or as the timed CPU baseline, never as the
product.  The product path (``episodic-transformer-memory-ppo_b200/``) never imports it and has no CPU
fallback: it raises if the CUDA library is missing.
"""
