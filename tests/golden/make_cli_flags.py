"""Golden list of the reference's command-line surface, extracted STATICALLY (ast) from /root/reference/code - the
reference cannot be imported (TensorFlow 1.12).  Writes tests/golden/cli_flags.json: per defining function the
ordered `add_argument` calls with name, default, type, action, choices, nargs.  Run here (the reference is mounted
read-only in the build container); the JSON travels, tests/test_cli_surface.py checks wlseg/settings.py against it.

usage: python tests/golden/make_cli_flags.py [/root/reference/code]
"""
import ast
import json
import os
import sys

SOURCES = [
    ('utils/utils.py', None),
    ('models/resnet50_extended_model_hierarchical.py', {'add_model_arguments'}),
    ('predict.py', None),
    ('evaluate.py', None),
    ('train.py', None),
    ('input_pipelines/cityscapes/input_cityscapes.py', None),
    ('input_pipelines/dataset_agnostic/dataset_agnostic_predict_input.py', None),
]


def literal(node):
  if node is None:
    return None
  try:
    return ast.literal_eval(node)
  except (ValueError, SyntaxError):
    return {'expr': ast.unparse(node)}


def extract(path, only):
  tree = ast.parse(open(path).read())
  out = {}
  for fn in ast.walk(tree):
    if not isinstance(fn, ast.FunctionDef) or (only and fn.name not in only):
      continue
    calls = []
    for node in ast.walk(fn):
      if isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) and node.func.attr == 'add_argument':
        if not node.args:
          continue
        name = literal(node.args[0])
        kw = {k.arg: k.value for k in node.keywords}
        t = kw.get('type')
        calls.append({'name': name, 'line': node.lineno,
                      'default': literal(kw.get('default')), 'has_default': 'default' in kw,
                      'type': None if t is None else ast.unparse(t), 'action': literal(kw.get('action')),
                      'choices': literal(kw.get('choices')), 'nargs': literal(kw.get('nargs'))})
    if calls:
      calls.sort(key=lambda c: c['line'])
      out[fn.name] = calls
  return out


def main():
  root = sys.argv[1] if len(sys.argv) > 1 else '/root/reference/code'
  golden = {}
  for rel, only in SOURCES:
    p = os.path.join(root, rel)
    if os.path.exists(p):
      got = extract(p, only)
      if got:
        golden[rel] = got
  dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'cli_flags.json')
  with open(dst, 'w') as fp:
    json.dump(golden, fp, indent=1, sort_keys=True)
  print(dst, {k: {f: len(v) for f, v in d.items()} for k, d in golden.items()})


if __name__ == '__main__':
  main()
