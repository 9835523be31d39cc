#!/usr/bin/env python
"""evaluate.py <log_dir> <Neval> <training_problem_def_path> <tfrecords_path> <per_pixel_dataset_name>
[flags] -- same surface as the reference's code/evaluate.py:25-83 (which is disabled upstream by a
`raise NotImplementedError`; enabled here).  Writes all_metrics.txt / all_metrics.p."""
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(_ROOT, 'iv2019-boosting-semantic-segmentation-with-weak-labels_b200'))

from wlseg.cli import evaluate_main  # noqa: E402

if __name__ == '__main__':
  evaluate_main(sys.argv[1:])
