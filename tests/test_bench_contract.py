"""bench.py's reference arm (`--impl reference`: the oracle port on the host cores) runs without a GPU, so its side of the
JSON contract is checked here on a tiny shape: one line on stdout, the base keys, `"impl": "reference"`, a
`cpu_baseline` describing the run, an `e2e` object without copies, the evaluation record nested under `"eval"`; under
torchrun only rank 0 prints, the other ranks exit 0 without work.  (The GPU arm's line is checked by the driver on the
B200; profiles/r2_bench_default_final.json holds the last one.)"""

import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
             'vs_baseline', 'dtype', 'data', 'config', 'e2e', 'gpu_launches', 'cpu_baseline', 'impl'}


def _run(extra_env=None, extra_args=()):
  env = dict(os.environ)
  env.update(extra_env or {})
  cmd = [sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0',
         '--height', '64', '--width', '96', '--batch', '2', *extra_args]
  return subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=300)


def _check_record(rec, metric, unit):
  assert BASE_KEYS <= set(rec), sorted(BASE_KEYS - set(rec))
  assert rec['impl'] == 'reference' and rec['metric'] == metric and rec['unit'] == unit
  assert rec['value'] > 0 and rec['higher_is_better'] is True and rec['vs_baseline'] is None and rec['gpu_launches'] == 0
  assert rec['data'] == 'synthetic' and rec['dtype'] == 'f32' and 'workload' in rec['config'] and 'model' not in rec['config']
  cpu = rec['cpu_baseline']
  assert set(cpu) >= {'value', 'unit', 'cores', 'kind', 'sample'} and cpu['kind'] == 'port' and cpu['value'] == rec['value']
  assert cpu['cores'] >= 1 and cpu['unit'] == unit
  e2e = rec['e2e']
  assert e2e == {'value': rec['value'], 'unit': unit, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}


def test_reference_arm_prints_one_contract_line():
  r = _run()
  assert r.returncode == 0, r.stderr[-2000:]
  lines = [l for l in r.stdout.splitlines() if l.strip()]
  assert len(lines) == 1, r.stdout
  rec = json.loads(lines[0])
  _check_record(rec, 'train_images_per_s', 'images/s')
  assert rec['steps'] == 1 and rec['warmup'] == 0 and rec['n_gpus'] == 1
  assert '2 strong + 0 bbox + 0 image-level images/GPU' in rec['config']['workload'] and '64x96' in rec['config']['workload']
  _check_record(rec['eval'], 'eval_mpix_per_s', 'Mpix/s')


def test_reference_arm_other_ranks_exit_without_work():
  r = _run({'RANK': '1', 'LOCAL_RANK': '1', 'WORLD_SIZE': '2'}, ('--gpus', '2'))
  assert r.returncode == 0 and r.stdout.strip() == '', (r.stdout, r.stderr[-500:])
