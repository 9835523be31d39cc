"""Digests of the reference's problem definitions:  tests/golden/problem_def_digests.json.

    python tests/golden/make_problem_def_digests.py       # needs /root/reference; run in the build container

code/problem_definitions/{cityscapes,vistas}/problem01.json are the files train.py / evaluate.py / predict.py read
(`training_problem_def_path`).  wlseg/problem_defs.py REGENERATES them from the public Cityscapes / Mapillary Vistas label
tables instead of carrying copies; this script stores, per field, the SHA-256 of the reference's value (canonical JSON),
so that tests/test_reference_fixtures.py can assert field-for-field equality without the reference's files travelling.
The free-text `comments` field is not compared.
"""

import hashlib
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get('WLSEG_REFERENCE', '/root/reference/code')
OUT = os.path.join(HERE, 'problem_def_digests.json')


def digest(value):
  return hashlib.sha256(json.dumps(value, sort_keys=True, separators=(',', ':')).encode('utf-8')).hexdigest()


def main():
  out = {}
  for dataset in ('cityscapes', 'vistas'):
    with open(os.path.join(REF, 'problem_definitions', dataset, 'problem01.json')) as fp:
      pd = json.load(fp)
    out[dataset] = {k: digest(v) for k, v in sorted(pd.items()) if k != 'comments'}
  with open(OUT, 'w') as fp:
    json.dump(out, fp, indent=1, sort_keys=True)
  print('wrote', OUT, {k: sorted(v) for k, v in out.items()})


if __name__ == '__main__':
  main()
