"""CPU oracle for the population-dynamics hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and there only as the
checker (or the timed CPU baseline), never as the thing shipped.  The product
package ``discrete_mean_field_game_b200`` must never import this package.

Parity status
-------------
* ``mfg_oracle`` (Dirichlet/softplus policy, mean-field step, closed-form
  rewards, quadratic-feature critic, TD(0) updates) is PINNED: it is checked in
  ``tests/test_oracle_golden.py`` against fixtures produced by running the
  reference's own ``mfg_ac2.py`` / ``mfg_ac.py`` / ``mfg_synthetic.py``
  (``oracle/make_golden.py``, fixtures in ``tests/golden/``).
* ``rnet_oracle`` (conv reward net, MaxEnt-IRL loss, TF-style Adam, z weights)
  is "parity unpinned" at the TensorFlow-1.x boundary: TensorFlow is not
  installable here and the reference records no expected values for it.  It is
  a float64 restatement of ``networks.py`` + ``ac_irl.py:382-418`` cross-checked
  by finite differences and by an independent Dirichlet-pdf formula
  (``test_acirl.py:15-30``).
"""
