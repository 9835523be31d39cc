"""Problem definitions (label-id -> class-id maps, class names, palettes).

The reference reads these from `problem_definitions/{cityscapes,vistas}/problem01.json`
(code/system_factory.py:79-97; fields lids2cids, cids2labels, cids2colors, cids2lids).  The
drop-in reads the same JSON schema from whatever path the user passes.  For the synthetic
benchmarks and the tests this module *generates* equivalent files from the public dataset
label tables (Cityscapes `labels.py` id/trainId table; Mapillary Vistas v1.2 `config.json`
label order), so that no reference file has to be present at run time.
"""

import json
import os

# (name, label id, train id or -1, colour) -- Cityscapes label table
_CITYSCAPES = [
    ('unlabeled', 0, -1, (0, 0, 0)), ('ego vehicle', 1, -1, (0, 0, 0)), ('rectification border', 2, -1, (0, 0, 0)),
    ('out of roi', 3, -1, (0, 0, 0)), ('static', 4, -1, (0, 0, 0)), ('dynamic', 5, -1, (111, 74, 0)),
    ('ground', 6, -1, (81, 0, 81)), ('road', 7, 0, (128, 64, 128)), ('sidewalk', 8, 1, (244, 35, 232)),
    ('parking', 9, -1, (250, 170, 160)), ('rail track', 10, -1, (230, 150, 140)), ('building', 11, 2, (70, 70, 70)),
    ('wall', 12, 3, (102, 102, 156)), ('fence', 13, 4, (190, 153, 153)), ('guard rail', 14, -1, (180, 165, 180)),
    ('bridge', 15, -1, (150, 100, 100)), ('tunnel', 16, -1, (150, 120, 90)), ('pole', 17, 5, (153, 153, 153)),
    ('polegroup', 18, -1, (153, 153, 153)), ('traffic light', 19, 6, (250, 170, 30)),
    ('traffic sign', 20, 7, (220, 220, 0)), ('vegetation', 21, 8, (107, 142, 35)), ('terrain', 22, 9, (152, 251, 152)),
    ('sky', 23, 10, (70, 130, 180)), ('person', 24, 11, (220, 20, 60)), ('rider', 25, 12, (255, 0, 0)),
    ('car', 26, 13, (0, 0, 142)), ('truck', 27, 14, (0, 0, 70)), ('bus', 28, 15, (0, 60, 100)),
    ('caravan', 29, -1, (0, 0, 90)), ('trailer', 30, -1, (0, 0, 110)), ('train', 31, 16, (0, 80, 100)),
    ('motorcycle', 32, 17, (0, 0, 230)), ('bicycle', 33, 18, (119, 11, 32)),
]

# Mapillary Vistas v1.2 labels in config.json order (name, colour); the last one is the void class
_VISTAS = [
    ('Bird', (165, 42, 42)), ('Ground Animal', (0, 192, 0)), ('Curb', (196, 196, 196)), ('Fence', (190, 153, 153)),
    ('Guard Rail', (180, 165, 180)), ('Barrier', (102, 102, 156)), ('Wall', (102, 102, 156)),
    ('Bike Lane', (128, 64, 255)), ('Crosswalk - Plain', (140, 140, 200)), ('Curb Cut', (170, 170, 170)),
    ('Parking', (250, 170, 160)), ('Pedestrian Area', (96, 96, 96)), ('Rail Track', (230, 150, 140)),
    ('Road', (128, 64, 128)), ('Service Lane', (110, 110, 110)), ('Sidewalk', (244, 35, 232)),
    ('Bridge', (150, 100, 100)), ('Building', (70, 70, 70)), ('Tunnel', (150, 120, 90)), ('Person', (220, 20, 60)),
    ('Bicyclist', (255, 0, 0)), ('Motorcyclist', (255, 0, 0)), ('Other Rider', (255, 0, 0)),
    ('Lane Marking - Crosswalk', (200, 128, 128)), ('Lane Marking - General', (255, 255, 255)),
    ('Mountain', (64, 170, 64)), ('Sand', (128, 64, 64)), ('Sky', (70, 130, 180)), ('Snow', (255, 255, 255)),
    ('Terrain', (152, 251, 152)), ('Vegetation', (107, 142, 35)), ('Water', (0, 170, 30)), ('Banner', (255, 255, 128)),
    ('Bench', (250, 0, 30)), ('Bike Rack', (0, 0, 0)), ('Billboard', (220, 220, 220)), ('Catch Basin', (170, 170, 170)),
    ('CCTV Camera', (222, 40, 40)), ('Fire Hydrant', (100, 170, 30)), ('Junction Box', (40, 40, 40)),
    ('Mailbox', (33, 33, 33)), ('Manhole', (170, 170, 170)), ('Phone Booth', (0, 0, 142)), ('Pothole', (170, 170, 170)),
    ('Street Light', (210, 170, 100)), ('Pole', (153, 153, 153)), ('Traffic Sign Frame', (128, 128, 128)),
    ('Utility Pole', (0, 0, 142)), ('Traffic Light', (250, 170, 30)), ('Traffic Sign (Back)', (192, 192, 192)),
    ('Traffic Sign (Front)', (220, 220, 0)), ('Trash Can', (180, 165, 180)), ('Bicycle', (119, 11, 32)),
    ('Boat', (0, 0, 142)), ('Bus', (0, 60, 100)), ('Car', (0, 0, 142)), ('Caravan', (0, 0, 90)),
    ('Motorcycle', (0, 0, 230)), ('On Rails', (0, 80, 100)), ('Other Vehicle', (128, 64, 64)), ('Trailer', (0, 0, 110)),
    ('Truck', (0, 0, 70)), ('Wheeled Slow', (0, 0, 192)), ('Car Mount', (32, 32, 32)), ('Ego Vehicle', (0, 0, 0)),
    ('Unlabeled', (0, 0, 0)),
]


def cityscapes():
  rows = sorted(_CITYSCAPES, key=lambda r: r[1])
  lids2cids = [r[2] for r in rows]
  n = max(lids2cids) + 1
  by_cid = {r[2]: r for r in rows if r[2] >= 0}
  return {
      'version': 2.0,
      'comments': 'cityscapes 19 train classes + void; void is -1 in lids2cids and the last class id internally',
      'lids2cids': lids2cids,
      'cids2labels': [by_cid[c][0] for c in range(n)] + ['void'],
      'cids2colors': [list(by_cid[c][3]) for c in range(n)] + [[0, 0, 0]],
      'cids2lids': [by_cid[c][1] for c in range(n)] + [0],
  }


def vistas():
  n = len(_VISTAS) - 1
  return {
      'version': 2.0,
      'comments': 'mapillary vistas 65 classes + void (Unlabeled); void is -1 and the last class id internally',
      'lids2cids': list(range(n)) + [-1],
      'cids2labels': [r[0] for r in _VISTAS],
      'cids2colors': [list(r[1]) for r in _VISTAS],
      'cids2lids': list(range(n)) + [-1],
  }


GENERATORS = {'cityscapes': cityscapes, 'vistas': vistas}


def default_path(dataset):
  """Path of the generated problem definition of `dataset` (written on first use; the files are
  build artefacts, not sources)."""
  here = os.path.dirname(os.path.abspath(__file__))
  path = os.path.join(here, 'problem_definitions', dataset, 'problem01.json')
  if not os.path.exists(path):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, 'w') as fp:
      json.dump(GENERATORS[dataset](), fp)
      fp.write('\n')
  return path


def write_all(root=None):
  """(Re)generate problem_definitions/<dataset>/problem01.json next to this module."""
  paths = []
  for name, gen in GENERATORS.items():
    path = default_path(name) if root is None else os.path.join(root, name, 'problem01.json')
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, 'w') as fp:
      json.dump(gen(), fp)
      fp.write('\n')
    paths.append(path)
  return paths


def load(path):
  with open(path, 'r') as fp:
    return json.load(fp)


if __name__ == '__main__':
  print('\n'.join(write_all()))
