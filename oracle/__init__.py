"""CPU oracle for the wlseg hot path -- TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (PyTorch fp32 on CPU for the floating-point
graph, numpy / plain C for the integer work) of the algorithm that the reference
(pmeletis/IV2019-boosting-semantic-segmentation-with-weak-labels, `code/`)
expresses through TensorFlow 1.12 ops.  Every function cites the reference
file:line it follows.

PARITY STATUS.  The arithmetic of this path lives in the un-vendored,
uninstallable third-party dependency `tensorflow==1.12.0`
(`code/requirements.txt:9`), and none of the reference's own tests exercises the
model, the loss, the estimator or the metrics.
PINNED (round 2) by golden vectors produced by RUNNING THE REFERENCE'S OWN PYTHON
over an emulation of the TF-1.12 calls it makes (tests/golden/tf_shim,
tests/golden/make_reference_fixtures.py -> tests/golden/reference_run.npz,
checked in tests/test_reference_fixtures.py): losses.py, weak_labels.py,
metrics.py, optimizer.py, preprocess.py and the remap / resize / void helpers.
PARITY UNPINNED for network.py / tfops.py (the ResNet-50 body is
`tf.contrib.slim`'s `resnet_v1_50`, un-vendored: restated from the published
behaviour of TF 1.12, pinned only by the worked examples the reference carries
in its comments, SURVEY.md section 8c, `tests/test_oracle_known_answers.py`).
The shim is itself our reading of TF's op semantics, so "pinned" means: the
reference's own call graph, constants, masks, tables and reductions are the ones
that ran - not that TensorFlow's kernels did.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import this package, and only as the checker.  The
product (`wlseg`) never imports it and has no CPU fallback.
"""
