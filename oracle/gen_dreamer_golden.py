"""Writes tests/golden/dreamer_tiny.npz: what the fp32 restatement of the reference's dreamerv3
update (oracle/dreamer_oracle.py) produces on a seeded tiny batch (TEST INFRA).

The reference itself cannot run here (no JAX), so this golden does not pin the oracle to the
reference -- it pins the oracle to ITSELF over time: `tests/test_dreamer_oracle.py` fails when
an edit to the oracle changes its numbers, and the GPU suite compares the product with the
same committed values on machines where the oracle and the product run side by side.
    python -m oracle.gen_dreamer_golden
"""
import pathlib
import sys

import numpy as np
import torch

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / 'tests'))
from oracle import dreamer_oracle as do     # noqa: E402

OUT = ROOT / 'tests' / 'golden' / 'dreamer_tiny.npz'
SPEC = dict(seed=0, B=3, T=6, steps=2)
PROBES = ('dyn/dyngru/kernel', 'dyn/obs0/kernel', 'enc/cnn1/kernel', 'dec/conv0/kernel',
          'pol/head/action/logits/kernel', 'val/head/logits/kernel', 'dyn/dynhid0norm/scale')


def run():
  """The seeded scenario: `steps` updates from freshly initialised parameters."""
  import dreamer_cases as cases
  ocfg = do.tiny_config()
  oracle = do.Dreamer(ocfg, do.init_params(ocfg, SPEC['seed'], outscale_override=1.0))
  out = {}
  for it in range(SPEC['steps']):
    data = cases.batch(ocfg, SPEC['B'], SPEC['T'], seed=10 + it)
    noise = do.make_noise(ocfg, SPEC['B'], SPEC['T'], seed=it)
    carry, outs, mets, grads, oo = oracle.train(data, noise)
    out[f's{it}/loss'] = mets['loss'].numpy()
    for k, v in oo['losses'].items():
      out[f's{it}/loss_{k}'] = v.detach().mean().numpy()
    out[f's{it}/index'] = oo['feat']['stoch'].argmax(-1).numpy().astype(np.int8)
    out[f's{it}/imgact'] = oo['imgact'].numpy().astype(np.int8)
    out[f's{it}/deter_last'] = carry['deter'].numpy()
    for k in PROBES:
      out[f's{it}/gradnorm/{k}'] = grads[k].double().norm().numpy()
      out[f's{it}/paramsum/{k}'] = oracle.p[k].double().sum().numpy()
  return out


if __name__ == '__main__':
  np.savez_compressed(OUT, **run())
  print(OUT, OUT.stat().st_size)
