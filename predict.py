#!/usr/bin/env python
"""predict.py <log_dir> <training_problem_def_path> <predict_dir> <per_pixel_dataset_name> [flags]
-- same surface as the reference's code/predict.py:22-220.  Plotting / export (matplotlib, PIL) is
out of scope; timing per image is printed as upstream."""
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(_ROOT, 'iv2019-boosting-semantic-segmentation-with-weak-labels_b200'))

from wlseg.cli import predict_main  # noqa: E402

if __name__ == '__main__':
  predict_main(sys.argv[1:])
