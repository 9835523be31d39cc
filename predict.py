#!/usr/bin/env python
"""predict.py <log_dir> <training_problem_def_path> <predict_dir> <per_pixel_dataset_name> [flags]
-- same surface as the reference's code/predict.py:22-220, including the PNG exports of :137-164
(--export_lids_images, --export_color_decisions, --export_overlapped_color_decisions into --results_dir;
wlseg/cli.py export_outputs).  Only the live matplotlib plotting of :40-135 is not built; timing per image is
printed as upstream."""
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(_ROOT, 'iv2019-boosting-semantic-segmentation-with-weak-labels_b200'))

from wlseg.cli import predict_main  # noqa: E402

if __name__ == '__main__':
  predict_main(sys.argv[1:])
