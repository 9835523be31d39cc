"""Build libwlseg.so (hand-written CUDA for sm_100a behind the C ABI of include/wlseg.h).

nvcc cross-compiles without a GPU; the .so is written IN-TREE (wlseg/lib/) so that it travels to
the GPU box with the repo snapshot.  No torch / pybind dependency: the library is plain C ABI.
"""

import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIBDIR = os.path.join(HERE, 'wlseg', 'lib')
LIB = os.path.join(LIBDIR, 'libwlseg.so')
SOURCES = ['abi.cu', 'confmat.cu', 'head.cu', 'loss.cu', 'pool.cu', 'bn.cu', 'optim.cu', 'conv_direct.cu',
           'conv_igemm_sm100.cu', 'conv_wgrad_sm100.cu', 'transform.cu', 'weak_labels.cu', 'psp.cu', 'postproc.cu', 'gn.cu', 'preproc.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '--expt-relaxed-constexpr', '-Xcompiler', '-fPIC',
              '-Xptxas', '-v']


def _nvcc():
  for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
    if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
      return cand
  raise RuntimeError('nvcc not found')


def _digest():
  h = hashlib.sha256()
  files = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))]
  files.append(os.path.join(HERE, '..', 'include', 'wlseg.h'))
  files.append(os.path.abspath(__file__))
  for f in files:
    with open(f, 'rb') as fp:
      h.update(fp.read())
  return h.hexdigest()


def build(force=False, verbose=False):
  """Compile every .cu to an object and link libwlseg.so; skipped when sources are unchanged."""
  os.makedirs(LIBDIR, exist_ok=True)
  stamp = os.path.join(LIBDIR, 'libwlseg.sha256')
  dig = _digest()
  if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read().strip() == dig:
    return LIB
  nvcc = _nvcc()
  objdir = os.path.join(HERE, 'build')
  os.makedirs(objdir, exist_ok=True)
  procs = []
  for src in SOURCES:
    obj = os.path.join(objdir, src.replace('.cu', '.o'))
    cmd = [nvcc] + NVCC_FLAGS + ['-c', os.path.join(CSRC, src), '-o', obj]
    procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
  objs = []
  log = []
  for src, obj, p in procs:
    out, _ = p.communicate()
    log.append(f'== {src}\n{out}')
    if p.returncode != 0:
      sys.stderr.write(out)
      raise RuntimeError(f'nvcc failed on {src}')
    objs.append(obj)
  with open(os.path.join(objdir, 'ptxas.log'), 'w') as fp:
    fp.write('\n'.join(log))
  if verbose:
    print('\n'.join(log))
  cmd = [nvcc, '-shared', '-o', LIB] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a', '-lcudart_static',
                                              '-Xcompiler', '-fPIC']
  r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
  if r.returncode != 0:
    sys.stderr.write(r.stdout)
    raise RuntimeError('link failed')
  with open(stamp, 'w') as fp:
    fp.write(dig)
  return LIB


if __name__ == '__main__':
  print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
