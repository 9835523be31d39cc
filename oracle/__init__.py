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
over an emulation of the TF-1.12 / tf.contrib.slim calls it makes
(tests/golden/tf_shim): tests/golden/reference_run.npz (losses.py,
weak_labels.py, metrics.py, optimizer.py, preprocess.py, the remap / resize /
void helpers) and tests/golden/reference_model_run.npz (network.py: the
reference's `model()` + `feature_extractor()` + `module_arg_scope()` executed
on 330-variable parameter dictionaries, five configurations incl. training-mode
batch norm, pyramid / field-of-view / hybrid upsampling and group norm) and
tests/golden/reference_driver_run.json (argument parsers, train.py / evaluate.py
overrides, SemanticSegmentation's derived settings), tests/golden/reference_train_run.npz
(train.py: `define_estimator` in TRAIN mode with the reference's model() for 3 / 2 optimizer
steps - losses, Momentum, EMA in UPDATE_OPS, moving statistics; plus the TRAIN graph's
global variables, warm-start map and saver with --init_ckpt_path),
tests/golden/reference_eval_run.npz (`define_estimator` in EVAL / PREDICT mode: cid map,
nearest resize to the label size, streaming confusion matrix over batches, restore names)
and tests/golden/reference_xreplica_run.npz (tfops.cross_replica_batch_norm: the reference's
CrossReplicaBatchNormalization._fused_batch_norm over two emulated towers) - checked in
tests/test_reference_fixtures.py (oracle) and tests/test_gpu_reference_fixtures.py (CUDA path,
no oracle in between).
What stays RESTATED: TensorFlow's and slim's own internals (un-vendored): the
shim is our reading of the op semantics (SAME padding split, fused batch norm,
resize_bilinear align_corners, resnet_v1's output-stride bookkeeping, variable
scoping rules ...), so "pinned" means: the reference's own call graph, wiring,
constants, masks, tables and reductions are the ones that ran - not that
TensorFlow's kernels did.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import this package, and only as the checker.  The
product (`wlseg`) never imports it and has no CPU fallback.
"""
