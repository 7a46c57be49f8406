"""torchrun worker of tests/test_gpu_exchange.py: every rank trains a tiny bf16 dreamerv3 on ITS OWN
batch for a few updates, once with the bucketed exchange (own NCCL communicator, all-reduce +
optimiser per bucket under the backward pass) and once with the plain path (one
torch.distributed all_reduce + one fused optimiser launch); prints one JSON line per rank."""
import json
import os
import sys
import pathlib

import numpy as np
import torch
import torch.distributed as dist

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / 'tests'))
from embodied_b200 import dreamerv3, elements       # noqa: E402
from oracle import dreamer_oracle as do              # noqa: E402
import dreamer_cases as cases                        # noqa: E402


def main():
  rank, world, local = (int(os.environ[k]) for k in ('RANK', 'WORLD_SIZE', 'LOCAL_RANK'))
  torch.cuda.set_device(local)
  dist.init_process_group('nccl', device_id=torch.device('cuda', local))
  graph = sys.argv[1] if len(sys.argv) > 1 else 'off'
  ocfg = do.tiny_config()
  vals = do.init_params(ocfg, 0, outscale_override=1.0)
  S = elements.Space
  obs = {'image': S(np.uint8, ocfg.image), 'reward': S(np.float32), 'is_first': S(bool),
         'is_last': S(bool), 'is_terminal': S(bool)}
  act = {'reset': S(bool), 'action': S(np.int32, (), 0, ocfg.actions)}
  agents = {}
  for name, buckets in (('bucketed', True), ('plain', False)):
    cfg = cases.product_config(ocfg, 'bfloat16')
    cfg['graph'], cfg['grad_buckets'] = graph, buckets
    agents[name] = dreamerv3.Agent(obs, act, cfg, values={k: v.numpy() for k, v in vals.items()})
  assert agents['bucketed'].exchange is not None and agents['plain'].exchange is None
  assert agents['bucketed'].exchange.comm is not None or world == 1
  B, T = 3, 6
  norms = {k: [] for k in agents}
  for it in range(5):
    data = cases.to_device(cases.batch(ocfg, B, T, seed=100 * rank + it))
    noise = cases.to_device(do.make_noise(ocfg, B, T, seed=1000 * rank + it))
    for name, a in agents.items():
      carry = a.init_train(B)
      _, _, mets = a.train(carry, data, cases.clone(noise))
      norms[name].append(float(mets['opt/grad_norm']))
  torch.cuda.synchronize()
  a, b = agents['bucketed'].store, agents['plain'].store
  diff = float((a.master - b.master).double().norm() / b.master.double().norm())
  # every rank must hold the same parameters
  mine = a.master.clone()
  ref = mine.clone()
  dist.broadcast(ref, 0)
  across = float((mine - ref).abs().max())
  low_ok = bool(torch.equal(a.low, a.master.to(torch.bfloat16)))
  print(json.dumps({'rank': rank, 'world': world, 'rel_diff_vs_plain': diff, 'max_diff_across_ranks': across,
                    'low_in_step': low_ok, 'grad_norms': norms,
                    'graphs': len(agents['bucketed']._graphs), 'expected': agents['bucketed'].exchange.expected}),
        flush=True)
  dist.barrier()
  os._exit(0)


if __name__ == '__main__':
  main()
