"""The drop-in boundary at the script level: every `add_argument` the reference makes for train.py / evaluate.py /
predict.py (extracted statically into tests/golden/cli_flags.json by tests/golden/make_cli_flags.py) exists in
wlseg/settings.py with the same name, default, type, action, choices and nargs, and the positionals come in the
reference's order.  References: code/utils/utils.py:34-174, code/models/resnet50_extended_model_hierarchical.py:228-269,
code/predict.py:22-35,171-197, code/evaluate.py:25-33, input_cityscapes.py:311-319,
dataset_agnostic_predict_input.py:156-164 (the training input pipeline adds no flags: per_pixel_per_bbox_per_image.py:89-104).
"""

import argparse
import json
import os

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, 'golden', 'cli_flags.json')))
U = GOLD['utils/utils.py']
MODEL = GOLD['models/resnet50_extended_model_hierarchical.py']['add_model_arguments']
PPD = GOLD['predict.py']['main']   # the per_pixel_dataset_name positional each script adds itself

# the `add_argument` sequence each reference script performs, in call order
SCRIPTS = {
    'train': U['add_system_arguments'] + U['add_tf_arguments'] + U['add_train_arguments'] + MODEL,
    'eval': (U['add_system_arguments'] + U['add_tf_arguments'] + U['add_evaluate_arguments'] +
             GOLD['input_pipelines/cityscapes/input_cityscapes.py']['add_evaluate_input_pipeline_arguments'] + MODEL +
             GOLD['evaluate.py']['main']),
    'infer': (U['add_system_arguments'] + U['add_tf_arguments'] + U['add_inference_arguments'] +
              GOLD['input_pipelines/dataset_agnostic/dataset_agnostic_predict_input.py']['add_predict_input_pipeline_arguments'] +
              MODEL + PPD + GOLD['predict.py']['_add_predict_arguments']),
}
TYPES = {'int': int, 'float': float, 'str': str, None: None}


def _parser(mode):
  from wlseg import settings as wsettings
  ss = wsettings.build_parser(mode)
  return ss.argparser if hasattr(ss, 'argparser') else ss._parser


@pytest.mark.parametrize('mode', ['train', 'eval', 'infer'])
def test_every_reference_flag_exists_with_the_same_definition(mode):
  p = _parser(mode)
  by_name = {}
  for a in p._actions:
    for s in (a.option_strings or [a.dest]):
      by_name[s] = a
  missing, wrong = [], []
  for c in SCRIPTS[mode]:
    name = c['name']
    a = by_name.get(name)
    if a is None:
      missing.append(name)
      continue
    if c['action'] == 'store_true':
      if not isinstance(a, argparse._StoreTrueAction):
        wrong.append((name, 'action', type(a).__name__))
      continue
    if a.type is not TYPES.get(c['type'], 'unknown'):
      wrong.append((name, 'type', a.type, c['type']))
    if c['has_default'] and not isinstance(c['default'], dict) and a.default != c['default']:
      wrong.append((name, 'default', a.default, c['default']))
    if c['choices'] is not None and set(a.choices or ()) != set(c['choices']):
      wrong.append((name, 'choices', a.choices, c['choices']))
    if c['nargs'] is not None and a.nargs != c['nargs']:
      wrong.append((name, 'nargs', a.nargs, c['nargs']))
  assert not missing, f'{mode}: reference flags missing from wlseg/settings.py: {missing}'
  assert not wrong, f'{mode}: definitions differ from the reference: {wrong}'


@pytest.mark.parametrize('mode', ['train', 'eval', 'infer'])
def test_positionals_come_in_the_reference_order(mode):
  want = [c['name'] for c in SCRIPTS[mode] if not c['name'].startswith('-')]
  got = [a.dest for a in _parser(mode)._actions if not a.option_strings]
  assert got == want, (got, want)


def test_reference_example_command_lines_parse():
  """code/README.md:24-31 style invocations (evaluate with its five positionals, predict with the README's flags)."""
  st = _parser('eval').parse_args(['/logs/run1', '500', 'problem_definitions/cityscapes/problem01.json', '/data/valFine.tfrecords',
                                   'cityscapes', '--Nb', '4', '--restore_emas', '--psp_module'])
  assert (st.Neval, st.Nb, st.restore_emas, st.psp_module, st.per_pixel_dataset_name) == (500, 4, True, True, 'cityscapes')
  st = _parser('infer').parse_args(['/logs/run1', 'pd.json', '/images', 'vistas', '--psp_module', '--export_color_decisions',
                                    '--results_dir', '/tmp/out', '--upsampling_method', 'hybrid', '--norm_layer', 'group'])
  assert (st.predict_dir, st.export_color_decisions, st.upsampling_method, st.norm_layer) == ('/images', True, 'hybrid', 'group')
  st = _parser('train').parse_args(['/logs/new', 'cityscapes', '--learning_rate_values', '0.01', '0.005', '--distribute'])
  assert st.learning_rate_values == [0.01, 0.005] and st.learning_rate_boundaries == [8, 15, 17] and st.distribute
