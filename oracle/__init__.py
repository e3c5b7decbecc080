"""
oracle/ -- CPU restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package, and there only as the checker
(or as the timed CPU reference arm), never as the thing shipped.  Nothing under
``diffxpy_b200/`` imports it.

PARITY STATUS
-------------
* ``oracle.stats`` (Wald / chi^2 Wald / LRT / Welch t / Mann-Whitney / BH and the
  two-group result constructors) is PINNED: ``tests/golden/make_golden.py`` ran the
  reference's own ``diffxpy/stats/stats.py`` (loaded from /root/reference, it imports
  with numpy+scipy only) and scipy on seeded inputs and committed the vectors under
  ``tests/golden/``; ``tests/test_oracle_golden.py`` checks the restatement against them.
* ``oracle.nb_glm`` (the negative-binomial GLM fit) is **PARITY UNPINNED**: the arithmetic
  lives in the third-party package ``batchglm`` (requirement ``batchglm>=0.7.4``,
  /root/reference/requirements.txt:5, setup.py:22 -- lower bound only, not vendored, not
  installable here).  The restatement follows the published batchglm 0.7.4 numpy backend
  (``batchglm/train/numpy/base_glm/estimator.py``, ``batchglm/models/glm_nb``) as recalled in
  SURVEY.md section 3.5, anchored on diffxpy's own call sites
  (diffxpy/testing/tests.py:161, 177-191, 222-229, 242-248) and on the reference's
  statistical acceptance tests (diffxpy/unit_test/test_single_null.py etc.).
"""
