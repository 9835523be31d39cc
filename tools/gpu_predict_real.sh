#!/bin/bash
# predict.py on real image files (dataset-agnostic input): two PNGs of different sizes -> colour / label-id PNGs at
# the raw image size.
python - <<'PY'
import numpy as np, os
from PIL import Image
os.makedirs('/tmp/imgs/sub', exist_ok=True); os.makedirs('/tmp/imgs_out', exist_ok=True)
rng = np.random.default_rng(0)
Image.fromarray(rng.integers(0, 256, (300, 520, 3), dtype=np.uint8)).save('/tmp/imgs/a.png')
Image.fromarray(rng.integers(0, 256, (256, 384, 3), dtype=np.uint8)).save('/tmp/imgs/sub/b.jpg')
PY
PD=iv2019-boosting-semantic-segmentation-with-weak-labels_b200/wlseg/problem_definitions/cityscapes/problem01.json
python predict.py /tmp/nolog $PD /tmp/imgs cityscapes --height_feature_extractor 256 --width_feature_extractor 512 --export_color_decisions --export_lids_images --export_overlapped_color_decisions --results_dir /tmp/imgs_out 2>&1 | tail -4
python -c "
from PIL import Image; import glob
for f in sorted(glob.glob('/tmp/imgs_out/*')): print(f, Image.open(f).size, Image.open(f).mode)"
