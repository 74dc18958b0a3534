"""CPU oracle for the JaxABM hot path -- TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (NumPy, plus a small C helper under
``oracle/c``) of the semantics of the reference's data-parallel hot path:

* ``jaxabm/agent.py:92-177``  (``AgentCollection.init`` / ``update``)
* ``jaxabm/model.py:118-262`` (``Model.initialize`` / ``step`` / ``run``)
* the third-party arithmetic those call: ``jax.random`` (threefry2x32 key
  algebra) and ``jax.vmap`` broadcasting -- JAX is *not* vendored in the
  reference (``requirements.txt:8-9`` pins ``jax>=0.4.1``) and is not
  installed in this image, so its published algorithm is restated in
  ``oracle/jaxlike.py``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import it.  The product package
``jaxabm_b200`` never imports anything from here and has no CPU fallback.

PARITY PINNING STATUS
---------------------
* threefry2x32 / ``PRNGKey`` / ``split`` / ``uniform`` / ``normal`` (both stream layouts): pinned by
  the published Random123 vectors and by the draws JAX's own documentation prints
  (``tests/golden/threefry_kat.json``; ``tests/test_oracle.py::test_jax_documented_draws``).
* Model loop, key schedule, step ordering: pinned by the reference's own
  behavioural tests (ported in ``tests/test_oracle.py::test_model_contract``,
  ``::test_agent_collection_contract`` and ``tests/test_host.py::test_api_contract_without_device``) and by
  closed forms.
* Floating trajectories of the workload rules and the Schelling / SIR rules:
  **parity unpinned** -- the reference holds no golden vectors for them
   and JAX cannot be run here to generate any.
* ``oracle/sharded.py`` restates the multi-GPU decompositions (grid row bands, network node ranges)
  rank by rank; ``tests/test_oracle_sharded.py`` pins them on the plain oracle.
"""
