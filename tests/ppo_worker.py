"""torchrun worker of tests/test_gpu_ppo.py::test_two_ranks_*: every rank runs three ppo updates
(gradient mean + normaliser moments exchanged over NCCL, embodied/jax/opt.py:52-54,
utils.py:76-81) on its own batch ('own') or on the same batch ('same'); prints one JSON line."""
import json
import os
import pathlib
import sys

import torch
import torch.distributed as dist

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / 'tests'))
from embodied_b200 import ppo                        # noqa: E402
from oracle import ppo_oracle as po                  # noqa: E402
import ppo_cases as cases                            # noqa: E402


def main():
  rank, world, local = (int(os.environ[k]) for k in ('RANK', 'WORLD_SIZE', 'LOCAL_RANK'))
  torch.cuda.set_device(local)
  dist.init_process_group('nccl', device_id=torch.device('cuda', local))
  mode = sys.argv[1] if len(sys.argv) > 1 else 'own'
  obs, act = cases.dummy_spaces()
  ocfg = po.tiny_config(warmup=2)
  oracle, vals = cases.oracle_for(ocfg, obs, act)
  agent = ppo.Agent(obs, act, cases.product_config(ocfg), values={k: v.numpy() for k, v in vals.items()})
  assert agent.world == world
  B, T = 3, 8
  carry = agent.init_train(B)
  zeros = {k: torch.zeros(B, *v.shape, dtype=torch.int32 if v.discrete else torch.float32) for k, v in act.items()}
  ocarry = (oracle.initial(B), zeros)
  losses = []
  for step in range(3):
    seed = 20 + step + (100 * rank if mode == 'own' else 0)
    data = cases.batch(ocfg, obs, act, B, T, seed=seed)
    carry, _, mets = agent.train(carry, cases.to_device(data))
    losses.append(float(mets['loss']))
    if mode == 'same' and rank == 0:
      ocarry, _, _, _, _ = oracle.train(ocarry, data)
  flat = agent.store.master.clone()
  gathered = [torch.empty_like(flat) for _ in range(world)]
  dist.all_gather(gathered, flat)
  across = max(float((g - gathered[0]).abs().max()) for g in gathered)
  norms = torch.cat([agent.advnorm.vars, agent.valnorm.vars])
  ng = [torch.empty_like(norms) for _ in range(world)]
  dist.all_gather(ng, norms)
  row = dict(rank=rank, mode=mode, losses=losses, max_diff_across_ranks=across,
             norm_diff_across_ranks=max(float((g - ng[0]).abs().max()) for g in ng),
             checksum=float(flat.double().abs().sum()))
  if mode == 'same' and rank == 0:
    worst = 0.0
    for k, v in oracle.p.items():
      d = float((agent.store.view('master', k).cpu() - v).abs().max())
      worst = max(worst, d / max(float(v.abs().max()), 1e-3))
    row['rel_diff_vs_oracle'] = worst
  sys.stdout.flush()
  for r in range(world):
    if r == rank:
      print(json.dumps(row), flush=True)
    dist.barrier()
  dist.destroy_process_group()


if __name__ == '__main__':
  main()
