#!/usr/bin/env python
"""train.py <log_dir> <per_pixel_dataset_name> [flags] -- same surface as the reference's
code/train.py:24-73 (positional order, flags and the hard overrides of `_add_extra_args`).
The input side is the on-device synthetic generator (`--synthetic` is implied: the reference's
TFRecord / Open Images pipelines are out of scope)."""
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(_ROOT, 'iv2019-boosting-semantic-segmentation-with-weak-labels_b200'))

from wlseg.cli import train_main  # noqa: E402

if __name__ == '__main__':
  train_main(sys.argv[1:])
