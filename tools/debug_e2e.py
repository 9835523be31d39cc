"""Where does the end-to-end evaluation number lose against the device-resident one?  Estimator.evaluate with host
batches at K = 20 / 60 steps per call (fixed per-call cost vs per-step cost), and with device-resident batches through
the same call (everything but the H2D copies)."""
import contextlib
import io
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'iv2019-boosting-semantic-segmentation-with-weak-labels_b200'))
import torch  # noqa: E402
from wlseg import problem_defs, settings as wsettings  # noqa: E402
from wlseg.system_factory import SemanticSegmentation  # noqa: E402

H, W, NB = 1024, 2048, 4
tmp = tempfile.mkdtemp()
ss = wsettings.build_parser(wsettings.EVAL)
st = wsettings.eval_extra_args(ss.parse_args([tmp, str(NB * 100), problem_defs.default_path('cityscapes'), 'synthetic', 'cityscapes',
                                              '--Nb', str(NB), '--height_feature_extractor', str(H), '--width_feature_extractor', str(W),
                                              '--synthetic']))
st.device, st.rank, st.world_size = 'cuda:0', 0, 1
g = torch.Generator().manual_seed(1)
host = []
for _ in range(2):
  host.append(({'proimages': (torch.rand((NB, H, W, 3), generator=g) * 2 - 1).pin_memory()},
               {'prolabels': torch.randint(0, 20, (NB, H, W), generator=g, dtype=torch.int32).pin_memory()}))
dev = [({'proimages': f['proimages'].cuda()}, {'prolabels': l['prolabels'].cuda()}) for f, l in host]
n = {'n': 5, 'src': host}


def input_fn(config, params):
  for i in range(n['n']):
    yield n['src'][i % 2]


system = SemanticSegmentation({'eval': input_fn}, None, st)
with contextlib.redirect_stdout(io.StringIO()):
  system.evaluate()
est = system.estimator
for src, name in ((host, 'host batches'), (dev, 'device batches')):
  for K in (20, 60):
    n['n'], n['src'] = K, src
    est.evaluate(input_fn(None, st), 20)
    ts = []
    for _ in range(3):
      torch.cuda.synchronize()
      t0 = time.perf_counter()
      est.evaluate(input_fn(None, st), 20)
      torch.cuda.synchronize()
      ts.append((time.perf_counter() - t0) * 1e3)
    print(f'{name:15s} K={K:3d}: ' + ' '.join(f'{t / K:7.3f}' for t in ts) + ' ms/step   (' + ' '.join(f'{t:7.1f}' for t in ts) + ' ms/call)', flush=True)
