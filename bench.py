"""Headline benchmark of the actor-learner hot path (BASELINE.json config 2).

One STEP = one iteration of ``embodied.run.train``'s inner loop at the reference
defaults (embodied/run/train.py:25-26,56-80): one Driver step over N=256 envs
(stack + normalise + policy + mask + replay append) followed by the
``train_ratio * N / (B*T)`` = 8 learner steps it triggers, each
``replay.sample(B=16, L=65)`` -> ``agent.train`` -> ``replay.update``.
Because the loop is serial, env-steps/s and learner-samples/s are locked by
``train_ratio``:  samples/s = 32 x env-steps/s.  ``value`` reports env-steps/s
and ``learner_samples_per_sec`` is printed beside it.

  value  observations already resident in HBM when the timed region starts
  e2e    the public API (Driver(...)(policy, steps=N) over host envs), pinned
         host staging -> H2D inside the timed region, masked actions D2H
  roofline   the dominant row-engine launch (replay gather), timed live with
         CUDA events on the launching stream
  cpu_baseline / --impl reference   the numpy restatement of the reference's own
         Driver + Replay + Consec code (oracle/host_oracle.py, pinned
         byte-for-byte against /root/reference), on the host cores

Launch: ``python bench.py --gpus N --steps K --warmup W`` (N>1 under torchrun,
one rank per GPU).  Each rank owns its own envs and replay shard (weak scaling,
no data-path collective: windows never span workers, embodied/core/replay.py:92).
"""
import argparse
import json
import os
import pathlib
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

B, T, PREFIX = 16, 64, 1            # dreamerv3/configs.yaml batch_size/length, replay_context
L = T + PREFIX
NENVS = 256
TRAIN_RATIO = 32
IMAGE = (64, 64, 3)
DETER, STOCH = 8192, (32, 64)       # size200m latents stored in replay (dreamerv3/agent.py:89-99)
CLASSES = 5
ROW_BYTES = 12288 + 32768 + 8192 + 20 + 4 + 4 + 3      # image, deter, stoch, stepid, reward, action, 3 flags
TRAINS_PER_STEP = TRAIN_RATIO * NENVS // (B * T)        # 8


def peaks():
  try:
    p = json.load(open(ROOT / 'MEASURED_PEAKS.json'))
    return float(p['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
  except Exception:
    return 6650.0, 'fallback (B200_PROFILING.md)'


# --------------------------------------------------------------------- clocks
class ClockSampler:
  QUERY = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,'
           'clocks_event_reasons.hw_thermal_slowdown,'
           'clocks_event_reasons.sw_thermal_slowdown,'
           'clocks_event_reasons.sw_power_cap')

  def __init__(self, index):
    self.index, self.rows, self.stop = index, [], threading.Event()
    self.thread = threading.Thread(target=self._run, daemon=True)

  def _run(self):
    while not self.stop.is_set():
      try:
        out = subprocess.run(
            ['nvidia-smi', f'--query-gpu={self.QUERY}', '--format=csv,noheader,nounits',
             '-i', str(self.index)], capture_output=True, text=True, timeout=5).stdout
        self.rows.append([x.strip() for x in out.strip().split(',')])
      except Exception:
        pass
      self.stop.wait(0.2)

  def __enter__(self):
    self.thread.start()
    return self

  def __exit__(self, *a):
    self.stop.set()
    self.thread.join(2)

  def summary(self):
    sm = [float(r[0]) for r in self.rows if len(r) == 6 and r[0].replace('.', '').isdigit()]
    mx = [float(r[1]) for r in self.rows if len(r) == 6 and r[1].replace('.', '').isdigit()]
    names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
    reasons = sorted({n for r in self.rows if len(r) == 6
                      for n, v in zip(names, r[2:]) if v.lower().startswith('active')})
    return {'sm_mhz': float(np.median(sm)) if sm else None,
            'sm_max_mhz': max(mx) if mx else None, 'reasons': reasons,
            'samples': len(sm)}


# ------------------------------------------------------------- the B200 arm
class FeedAgent:
  """Stand-in learner for the plumbing-only workload: latents come from a
  precomputed device pool, train() hands them back for Replay.update.  Keeps the
  Agent protocol (embodied/core/base.py:1-31) so Driver/Replay run unchanged."""

  device_obs = True

  def __init__(self, torch, n, seed):
    from embodied_b200 import elements
    self.torch = torch
    g = torch.Generator(device='cuda').manual_seed(seed)
    self.ext_space = {
        'dyn/deter': elements.Space(np.float32, (DETER,)),
        'dyn/stoch': elements.Space(np.float32, STOCH)}
    self.deter = torch.randn((4, n, DETER), generator=g, device='cuda')
    self.stoch = torch.randn((4, n, *STOCH), generator=g, device='cuda')
    self.action = torch.randint(0, CLASSES, (4, n), generator=g, device='cuda', dtype=torch.int32)
    self.t = 0

  def init_policy(self, n):
    return ()

  def policy(self, carry, obs, mode='train'):
    i = self.t % 4
    self.t += 1
    return carry, {'action': self.action[i]}, {
        'dyn/deter': self.deter[i], 'dyn/stoch': self.stoch[i]}

  def train(self, carry, data):
    out = {'replay': {'stepid': data['stepid'][:, PREFIX:],
                      'dyn/deter': data['dyn/deter'][:, PREFIX:],
                      'dyn/stoch': data['dyn/stoch'][:, PREFIX:]}}
    return carry, out, {}


def make_env(i):
  from embodied_b200.envs import synthetic
  return synthetic.SyntheticImage(i, size=IMAGE, classes=CLASSES, length=500)


class Loop:
  """Config-2 wiring (dreamerv3/main.py:183-272): Replay(length=65, online,
  chunksize 1024) + Stateless -> Consec(length 64, consec 1, prefix 1)."""

  def __init__(self, torch, rank, capacity):
    import embodied_b200 as embodied
    from embodied_b200 import _lib
    self.torch, self.lib, self.embodied = torch, _lib, embodied
    self.replay = embodied.Replay(
        L, capacity, chunksize=1024, online=True, seed=0,
        staging_rows=NENVS, workers=NENVS)
    self.agent = FeedAgent(torch, NENVS, seed=rank)
    base = embodied.streams.Stateless(self.replay.sample, B, 'train')
    self.stream = iter(embodied.streams.Consec(
        base, length=T, consec=1, prefix=PREFIX, strict=True, contiguous=True))
    self.driver = embodied.Driver(
        [(lambda i=i: make_env(rank * NENVS + i)) for i in range(NENVS)],
        parallel=False, fetch_outs=False)
    self.driver.on_step(self.replay.add)
    self.driver.on_batch(lambda trans, n: self.learn())
    self.driver.reset(self.agent.init_policy)
    self.carry = None
    self.learner_on = False
    self.result = None

  def learn(self):
    if not self.learner_on:
      return
    for _ in range(TRAINS_PER_STEP):
      batch = next(self.stream)
      self.carry, outs, _ = self.agent.train(self.carry, batch)
      self.replay.update(outs['replay'])
      self.result = batch['reward']

  def step_e2e(self):
    """The user-facing call: host envs -> pinned staging -> H2D -> kernels."""
    self.driver(self.agent.policy, steps=NENVS)
    return self.result.sum(1)[:1].cpu()      # D2H read of the step's result

  def make_resident(self):
    """Pre-stage the observations of a Driver step in HBM (for `value`)."""
    torch = self.torch
    g = torch.Generator(device='cuda').manual_seed(7)
    self.res = {
        'image': torch.randint(0, 256, (NENVS, *IMAGE), generator=g, device='cuda', dtype=torch.uint8),
        'reward': torch.randn(NENVS, generator=g, device='cuda'),
        'is_first': torch.zeros(NENVS, dtype=torch.bool, device='cuda'),
        'is_last': torch.zeros(NENVS, dtype=torch.bool, device='cuda'),
        'is_terminal': torch.zeros(NENVS, dtype=torch.bool, device='cuda')}

  def step_resident(self):
    """Same work with the N observations already on the device: policy, append
    (one launch), then the learner steps."""
    _, acts, outs = self.agent.policy((), self.res)
    self.replay.add_batch({**self.res, **acts, **outs})
    self.learn()


def run_b200(args):
  import torch
  import torch.distributed as dist
  rank = int(os.environ.get('RANK', 0))
  world = int(os.environ.get('WORLD_SIZE', 1))
  local = int(os.environ.get('LOCAL_RANK', 0))
  assert torch.cuda.is_available(), 'bench.py needs a CUDA device (no CPU fallback)'
  torch.cuda.set_device(local)
  if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
  assert world == args.gpus, (world, args.gpus)
  from embodied_b200 import _lib
  from embodied_b200.core import store as storelib
  peak, peak_src = peaks()

  capacity = int(args.capacity)
  loop = Loop(torch, rank, capacity)
  flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')

  # prefill through the public API until windows exist
  while len(loop.replay) < 4 * B * L:
    loop.driver(loop.agent.policy, steps=NENVS)
  loop.learner_on = True
  loop.make_resident()

  def barrier():
    torch.cuda.synchronize()
    if world > 1:
      dist.barrier()
      torch.cuda.synchronize()

  def timed(fn, steps, warmup, profile=False):
    for _ in range(warmup):
      fn()
    barrier()
    storelib.PROFILE = [] if profile else None
    launches0 = _lib.launch_count()
    total = 0.0
    for _ in range(steps):
      flush.zero_()                                  # L2 flush, untimed
      a = torch.cuda.Event(enable_timing=True)
      b = torch.cuda.Event(enable_timing=True)
      torch.cuda.synchronize()
      a.record()
      fn()
      b.record()
      torch.cuda.synchronize()
      total += a.elapsed_time(b) * 1e-3
    barrier()
    launches = _lib.launch_count() - launches0
    prof, storelib.PROFILE = storelib.PROFILE, None
    t = torch.tensor([total], device='cuda', dtype=torch.float64)
    if world > 1:
      dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()), launches, prof

  with ClockSampler(local) as clocks:
    t_dev, launches, prof = timed(loop.step_resident, args.steps, args.warmup, profile=True)
    t_e2e, _, _ = timed(loop.step_e2e, args.steps, args.warmup)
  clk = clocks.summary()

  # dominant kernel: the replay gather launch (2 x B x L x row bytes per launch)
  gather_us = [a.elapsed_time(b) * 1e3 for kind, a, b in prof if kind == 'gather']
  gather_bytes = 2 * B * L * (ROW_BYTES + 4)        # + the int32 consec key
  g_us = float(np.mean(gather_us)) if gather_us else float('nan')
  achieved = gather_bytes / (g_us * 1e-6) / 1e9

  env_steps = world * NENVS * args.steps
  h2d = NENVS * (12288 + 4 + 3 + 20 + 8) + TRAINS_PER_STEP * (B * L * 8 + B * T * 8)
  d2h = NENVS * 4 + 4
  line = {
      'metric': 'env_steps_per_sec', 'value': env_steps / t_dev, 'unit': 'env steps/s',
      'learner_samples_per_sec': env_steps / t_dev * TRAIN_RATIO,
      'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
      'ms_per_step': t_dev / args.steps * 1e3, 'higher_is_better': True,
      'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u8', 'data': 'synthetic',
      'config': {
          'workload': 'config 2 PLUMBING ONLY: Driver(256 envs, 64x64x3 u8) + Replay(L=65, '
                      '53 299 B rows incl. f32 latents) append/sample(B=16)/update, '
                      f'{TRAINS_PER_STEP} learner iterations per Driver step (train_ratio 32); '
                      'NO model yet: latents come from a device pool',
          'envs_per_gpu': NENVS, 'batch': [B, T], 'replay_capacity_items': capacity,
          'cache': 'L2 flushed (256 MiB write) before every timed step; replay tables > L2'},
      'e2e': {'value': env_steps / t_e2e, 'unit': 'env steps/s',
              'ms_per_step': t_e2e / args.steps * 1e3,
              'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h},
      'gpu_launches': launches,
      'roofline': {'kernel': 'rows_kernel (emb_replay_gather, B=16 L=65)', 'bound': 'hbm',
                   'achieved': achieved, 'peak': peak, 'peak_source': peak_src,
                   'unit': 'GB/s', 'frac': achieved / peak, 'us_per_launch': g_us,
                   'launches_timed': len(gather_us), 'algorithmic_bytes': gather_bytes,
                   'traffic': 64950000},   # profiles/r01_rows_kernel_gather_B16_L65.md
      'clocks': clk,
  }
  if rank == 0:
    if world == 1 and not args.no_cpu:
      line['cpu_baseline'] = cpu_baseline(seconds=args.cpu_seconds)
    print(json.dumps(line), flush=True)
  if world > 1:
    dist.destroy_process_group()


# ------------------------------------------------------ the reference (CPU) arm
class OracleLoop:
  """The same iteration on the host with the numpy restatement of the
  reference's code (oracle/host_oracle.py).  bench.py is one of the few places
  allowed to execute oracle/; it is the thing compared against, never shipped."""

  def __init__(self, rank=0):
    import itertools
    from oracle import host_oracle as ho
    self.ho = ho
    envs = [make_env(rank * NENVS + i) for i in range(NENVS)]
    self.replay = ho.OracleReplay(L, None, 1024, True, 0, ids=itertools.count(1))
    self.driver = ho.OracleDriver(envs, envs[0].act_space)
    self.driver.callbacks.append(lambda row, w: self.replay.add(row, w))
    rng = np.random.default_rng(rank)
    self.deter = rng.standard_normal((4, NENVS, DETER), dtype=np.float32)
    self.stoch = rng.standard_normal((4, NENVS, *STOCH), dtype=np.float32)
    self.action = rng.integers(0, CLASSES, (4, NENVS)).astype(np.int32)
    self.t = 0
    self.learner_on = False

  def policy(self, carry, obs):
    i = self.t % 4
    self.t += 1
    self.ho.normalize_image(obs['image'])              # rssm.py:230 on the host
    return carry, {'action': self.action[i]}, {
        'dyn/deter': self.deter[i], 'dyn/stoch': self.stoch[i]}

  def step(self):
    self.driver.step(self.policy)
    if not self.learner_on:
      return
    for _ in range(TRAINS_PER_STEP):
      batch = self.ho.consec_view(self.replay.sample(B), T, 0, PREFIX)
      self.replay.update({k: batch[k][:, PREFIX:] for k in ('stepid', 'dyn/deter', 'dyn/stoch')})


def time_oracle(steps, warmup):
  loop = OracleLoop()
  while len(loop.replay) < 4 * B * L:
    loop.step()
  loop.learner_on = True
  for _ in range(warmup):
    loop.step()
  t0 = time.perf_counter()
  for _ in range(steps):
    loop.step()
  return time.perf_counter() - t0


def cpu_baseline(seconds=15.0):
  t1 = time_oracle(1, 1)
  steps = int(max(2, min(200, seconds / max(t1, 1e-3))))
  t = time_oracle(steps, 0)
  return {'value': NENVS * steps / t, 'unit': 'env steps/s', 'cores': 1, 'kind': 'port',
          'sample': f'{steps} iterations of the same config-2 loop (256 envs + {TRAINS_PER_STEP} '
                    'sample/update) with oracle/host_oracle.py, single thread as in '
                    'run.train debug mode (dreamerv3/configs.yaml:70)',
          'ms_per_step': t / steps * 1e3}


def run_reference(args):
  if int(os.environ.get('RANK', 0)) != 0:
    return
  steps = max(1, min(args.steps, 50))
  t = time_oracle(steps, min(args.warmup, 3))
  v = NENVS * steps / t
  print(json.dumps({
      'impl': 'reference', 'metric': 'env_steps_per_sec', 'value': v, 'unit': 'env steps/s',
      'learner_samples_per_sec': v * TRAIN_RATIO,
      'n_gpus': args.gpus, 'steps': steps, 'warmup': min(args.warmup, 3),
      'ms_per_step': t / steps * 1e3, 'higher_is_better': True, 'scaling': 'weak',
      'vs_baseline': None, 'dtype': 'u8', 'data': 'synthetic',
      'config': {'workload': 'config 2 PLUMBING ONLY (same as the b200 arm), numpy port of the '
                             'reference Driver+Replay+Consec on the host', 'envs': NENVS,
                 'batch': [B, T]},
      'cpu_baseline': {'value': v, 'unit': 'env steps/s', 'cores': 1, 'kind': 'port',
                       'sample': f'{steps} iterations'},
      'e2e': {'value': v, 'unit': 'env steps/s', 'h2d_bytes_per_step': 0,
              'd2h_bytes_per_step': 0}}), flush=True)


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=30)
  ap.add_argument('--warmup', type=int, default=5)
  ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
  ap.add_argument('--capacity', type=float, default=2e5)
  ap.add_argument('--cpu-seconds', type=float, default=15.0)
  ap.add_argument('--no-cpu', action='store_true')
  args = ap.parse_args()
  args.warmup = max(args.warmup, 3)
  if args.impl == 'reference':
    run_reference(args)
  else:
    run_b200(args)


if __name__ == '__main__':
  main()
