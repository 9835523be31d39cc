"""CPU oracle for the wlseg hot path -- TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (PyTorch fp32 on CPU for the floating-point
graph, numpy / plain C for the integer work) of the algorithm that the reference
(pmeletis/IV2019-boosting-semantic-segmentation-with-weak-labels, `code/`)
expresses through TensorFlow 1.12 ops.  Every function cites the reference
file:line it follows.

PARITY UNPINNED.  The arithmetic of this path lives in the un-vendored,
uninstallable third-party dependency `tensorflow==1.12.0`
(`code/requirements.txt:9`), and none of the reference's own tests exercises the
model, the loss, the estimator or the metrics.  The oracle
is therefore pinned only against the worked examples the reference carries in
its comments (`tests/test_oracle_known_answers.py`); the
TF-internal semantics are restated from the published behaviour of TF 1.12.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import this package, and only as the checker.  The
product (`wlseg`) never imports it and has no CPU fallback.
"""
