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
# dram__bytes_read.sum + dram__bytes_write.sum per launch from `ncu --set full` (profiles/)
NCU_TRAFFIC = {'rssm_fwd': 8863621000 + 194722816,      # profiles/r01_rssm_fwd_tma_kernel.md
               'rssm_bwd': 10104900000 + 137023232}     # profiles/r01_rssm_bwd_tma_kernel.md


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

  init_train = init_policy

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

  def __init__(self, torch, rank, capacity, agent='dreamerv3', size='size200m',
               dtype='bfloat16'):
    import embodied_b200 as embodied
    from embodied_b200 import _lib
    self.torch, self.lib, self.embodied = torch, _lib, embodied
    self.replay = embodied.Replay(
        L, capacity, chunksize=1024, online=True, seed=0,
        staging_rows=NENVS, workers=NENVS)
    if agent == 'feed':
      self.agent = FeedAgent(torch, NENVS, seed=rank)
    else:
      from embodied_b200 import dreamerv3
      env = make_env(0)
      self.agent = dreamerv3.Agent(
          env.obs_space, env.act_space,
          dreamerv3.config.make(size, compute_dtype=dtype, seed=0,
                                graph=os.environ.get('EMB_GRAPH', 'auto')))
    base = embodied.streams.Stateless(self.replay.sample, B, 'train')
    self.stream = iter(embodied.streams.Consec(
        base, length=T, consec=1, prefix=PREFIX, strict=True, contiguous=True))
    self.driver = embodied.Driver(
        [(lambda i=i: make_env(rank * NENVS + i)) for i in range(NENVS)],
        parallel=False, fetch_outs=False)
    self.driver.on_step(self.replay.add)
    self.driver.on_batch(lambda trans, n: self.learn())
    self.driver.reset(self.agent.init_policy)
    self.carry = self.agent.init_train(B) if agent != 'feed' else None
    self.pcarry = self.agent.init_policy(NENVS)
    self.learner_on = False
    self.result = None

  def learn(self):
    if not self.learner_on:
      return
    for _ in range(TRAINS_PER_STEP):
      batch = next(self.stream)
      self.carry, outs, mets = self.agent.train(self.carry, batch)
      self.replay.update(outs['replay'])
      self.result = mets.get('loss', batch['reward'][0, 0])

  def step_e2e(self):
    """The user-facing call: host envs -> pinned staging -> H2D -> kernels."""
    self.driver(self.agent.policy, steps=NENVS)
    return self.result.cpu()                  # D2H read of the step's result (loss)

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
    self.pcarry, acts, outs = self.agent.policy(self.pcarry, self.res)
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
  from embodied_b200.dreamerv3 import scan as scanlib
  peak, peak_src = peaks()

  capacity = int(args.capacity)
  scanlib.GRAPH_TIMERS = {}       # stopwatches recorded inside the captured train step
  loop = Loop(torch, rank, capacity, args.agent, args.size, args.dtype)
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
    scanlib.PROFILE = [] if profile else None
    launches0 = _lib.launch_count()
    total = 0.0
    graph_us = []
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
      if profile and getattr(loop.agent, '_graphs', None):
        for k, w in list(scanlib.GRAPH_TIMERS.items()):   # the step's last replay of the captured graph
          try:
            graph_us.append((k, w.ms() * 1e3))
          except RuntimeError:                       # never recorded (capture refused): eager events cover it
            scanlib.GRAPH_TIMERS.pop(k, None)
    barrier()
    launches = _lib.launch_count() - launches0
    prof, storelib.PROFILE = storelib.PROFILE, None
    prof = [(k, a.elapsed_time(b) * 1e3) for k, a, b in (prof or [])]
    if scanlib.PROFILE:
      prof += [(k, a.elapsed_time(b) * 1e3) for k, a, b, _ in scanlib.PROFILE]
    prof += graph_us
    scanlib.PROFILE = None
    t = torch.tensor([total], device='cuda', dtype=torch.float64)
    if world > 1:
      dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()), launches, prof

  with ClockSampler(local) as clocks:
    t_dev, launches, prof = timed(loop.step_resident, args.steps, args.warmup, profile=True)
    t_e2e, _, _ = timed(loop.step_e2e, args.steps, args.warmup)
  clk = clocks.summary()

  # per-kernel live timings (CUDA events on the launching stream, inside the timed region)
  def kernel_line(kind, name, nbytes, traffic):
    us = [t for k, t in prof if k == kind]
    if not us:
      return None
    t = float(np.mean(us))
    ach = nbytes / (t * 1e-6) / 1e9
    return {'kernel': name, 'bound': 'hbm', 'achieved': ach, 'peak': peak, 'peak_source': peak_src,
            'unit': 'GB/s', 'frac': ach / peak, 'us_per_launch': t, 'launches_timed': len(us),
            'algorithmic_bytes': nbytes, 'traffic': traffic}
  gather_bytes = 2 * B * L * (ROW_BYTES + 4)        # + the int32 consec key
  kernels = [kernel_line('gather', 'rows_kernel (emb_replay_gather, B=16 L=65)', gather_bytes,
                         64950000)]                  # profiles/r01_rows_kernel_gather_B16_L65.md
  if args.agent != 'feed':
    cfg = loop.agent.cfg
    Dg = cfg.deter // cfg.blocks
    # every in-scan weight once per step; the forward bf16 kernel hoists the action
    # rows of dynhid0 out of the scan (Dg + 2H input rows instead of Dg + 3H)
    fixed = cfg.deter * 2 * cfg.hidden + cfg.hidden * cfg.stoch * cfg.classes + cfg.deter * 3 * Dg
    esz = 2 if args.dtype == 'bfloat16' else 4
    wbytes = (fixed + cfg.deter * (Dg + 3 * cfg.hidden)) * esz * T
    wbytes_fwd = (fixed + cfg.deter * (Dg + (2 if args.dtype == 'bfloat16' else 3) * cfg.hidden)) * esz * T
    # DRAM traffic per launch from ncu (profiles/r01_rssm_*_kernel.md), bf16 size200m only
    known = args.dtype == 'bfloat16' and args.size == 'size200m'
    kernels.append(kernel_line('rssm_bwd', 'rssm_bwd_kernel (emb_rssm_observe_bwd, TMA weight ring, B=16 T=64)', wbytes_fwd,
                               NCU_TRAFFIC.get('rssm_bwd') if known else None))
    kernels.append(kernel_line('rssm_fwd', 'rssm_fwd_kernel (emb_rssm_observe_fwd, TMA weight ring, B=16 T=64)', wbytes_fwd,
                               NCU_TRAFFIC.get('rssm_fwd') if known else None))
  kernels = [k for k in kernels if k]
  # the roofline line is the hand-written kernel with the largest share of the step
  per_step = {'rows_kernel': TRAINS_PER_STEP, 'rssm_bwd_kernel': TRAINS_PER_STEP,
              'rssm_fwd_kernel': TRAINS_PER_STEP}
  share = lambda k: k['us_per_launch'] * per_step[k['kernel'].split()[0]] * args.steps
  roofline = max(kernels, key=share)
  for k in kernels:
    k['share_of_step'] = share(k) * 1e-6 / t_dev

  env_steps = world * NENVS * args.steps
  h2d = NENVS * (12288 + 4 + 3 + 20 + 8) + TRAINS_PER_STEP * (B * L * 8 + B * T * 8)
  d2h = NENVS * 4 + 4
  line = {
      'metric': 'env_steps_per_sec', 'value': env_steps / t_dev, 'unit': 'env steps/s',
      'learner_samples_per_sec': env_steps / t_dev * TRAIN_RATIO,
      'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
      'ms_per_step': t_dev / args.steps * 1e3, 'higher_is_better': True,
      'scaling': 'weak', 'vs_baseline': None,
      'dtype': 'u8' if args.agent == 'feed' else {'bfloat16': 'bf16', 'float32': 'f32'}[args.dtype],
      'data': 'synthetic',
      'config': {
          'workload': workload_name(args),
          'envs_per_gpu': NENVS, 'batch': [B, T], 'train_ratio': TRAIN_RATIO,
          'learner_steps_per_step': TRAINS_PER_STEP, 'replay_capacity_items': capacity,
          'parallelism': f'dp{world}: envs, replay shard and batch per rank; NCCL grad all-reduce',
          'cache': 'L2 flushed (256 MiB write) before every timed step; replay tables > L2'},
      'e2e': {'value': env_steps / t_e2e, 'unit': 'env steps/s',
              'ms_per_step': t_e2e / args.steps * 1e3,
              'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h},
      'gpu_launches': launches,
      'roofline': roofline, 'kernels': kernels,
      'clocks': clk,
  }
  if rank == 0:
    if world == 1 and not args.no_cpu:
      line['cpu_baseline'] = cpu_baseline(args)
    print(json.dumps(line), flush=True)
  if world > 1:
    # Captured CUDA graphs hold NCCL work: destroy_process_group() blocks on them.  All
    # ranks have finished (barrier) and rank 0 has printed; leave without the teardown.
    sys.stdout.flush()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    os._exit(0)


def workload_name(args):
  if args.agent == 'feed':
    return ('config 2 PLUMBING ONLY: Driver(256 envs, 64x64x3 u8) + Replay(L=65, 53 299 B rows '
            'incl. f32 latents) append/sample(B=16)/update; no model, latents from a device pool')
  return (f'config 2: dreamerv3 {args.size} on synthetic 64x64x3 image env, 256 envs, '
          'replay (B=16,T=64,L=65), one Driver step + 8 x (sample -> train -> update)')


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


def time_plumbing(steps, warmup):
  loop = OracleLoop()
  while len(loop.replay) < 4 * B * L:
    loop.step()
  loop.learner_on = True
  for _ in range(warmup):
    loop.step()
  t0 = time.perf_counter()
  for _ in range(steps):
    loop.step()
  return (time.perf_counter() - t0) / steps


class OracleLearner:
  """The fp32 torch-CPU restatement of the reference's dreamerv3 step
  (oracle/dreamer_oracle.py) at the benchmark's model size, all host threads."""

  def __init__(self, size):
    import torch
    from oracle import dreamer_oracle as do
    from embodied_b200.dreamerv3 import config as C
    self.torch, self.do = torch, do
    self.threads = os.cpu_count() or 1
    torch.set_num_threads(self.threads)
    self.cfg = do.default_config(**C.SIZES[size])
    self.model = do.Dreamer(self.cfg, do.init_params(self.cfg, 0))

  def batch(self, b):
    torch, cfg = self.torch, self.cfg
    g = torch.Generator().manual_seed(b)
    return {
        'image': torch.randint(0, 256, (b, L, *IMAGE), generator=g, dtype=torch.uint8),
        'reward': torch.randn(b, L, generator=g),
        'is_first': torch.zeros(b, L, dtype=torch.bool),
        'is_last': torch.zeros(b, L, dtype=torch.bool),
        'is_terminal': torch.zeros(b, L, dtype=torch.bool),
        'action': torch.randint(0, CLASSES, (b, L), generator=g, dtype=torch.int32),
        'dyn/deter': torch.zeros(b, L, cfg.deter),
        'dyn/stoch': torch.zeros(b, L, cfg.stoch, cfg.classes),
        'stepid': torch.zeros(b, L, 20, dtype=torch.uint8)}

  def time_train(self, b):
    data, noise = self.batch(b), self.do.make_noise(self.cfg, b, T, 0)
    t0 = time.perf_counter()
    self.model.train(data, noise)
    return time.perf_counter() - t0

  def time_policy(self, n):
    torch, cfg = self.torch, self.cfg
    carry = dict(deter=torch.zeros(n, cfg.deter), stoch=torch.zeros(n, cfg.stoch, cfg.classes),
                 action=torch.zeros(n, dtype=torch.int32))
    image = torch.randint(0, 256, (n, *IMAGE), dtype=torch.uint8)
    noise = dict(stoch=torch.zeros(n, cfg.stoch, cfg.classes), action=torch.zeros(n, cfg.actions))
    t0 = time.perf_counter()
    self.model.policy(carry, image, torch.zeros(n, dtype=torch.bool), noise)
    return time.perf_counter() - t0


def cpu_iteration(args, learner=None):
  """Seconds per benchmark step on the host, from a bounded sample: the
  Driver+Replay iteration is timed whole; agent.policy is timed on 64 of the 256
  envs (x4); agent.train on sub-batches B=1 and B=2 of the (16, 64) batch and
  extrapolated linearly to B=16 (the per-step weight traffic does not scale
  with B, so a plain x16 would overstate the CPU time)."""
  t_plumb = time_plumbing(2, 1)
  if args.agent == 'feed':
    return t_plumb, {'plumbing_s': t_plumb}, 1
  learner = learner or OracleLearner(args.size)
  t_pol = 4 * learner.time_policy(64)
  t1 = learner.time_train(1)
  t2 = learner.time_train(2)
  t16 = t1 + 15 * max(t2 - t1, 0.0)
  parts = {'plumbing_s': t_plumb, 'policy_256_s': t_pol, 'train_B1_s': t1, 'train_B2_s': t2,
           'train_B16_extrapolated_s': t16}
  return t_plumb + t_pol + TRAINS_PER_STEP * t16, parts, learner.threads


def cpu_baseline(args):
  t, parts, threads = cpu_iteration(args)
  return {'value': NENVS / t, 'unit': 'env steps/s', 'cores': threads, 'kind': 'port',
          'sample': 'one benchmark step assembled from a bounded sample: oracle Driver+Replay '
                    'iteration timed whole (1 thread, as in run.train debug mode); oracle '
                    'dreamerv3 policy on 64/256 envs x4; oracle train step at B=1 and B=2 '
                    f'(T=64) extrapolated linearly to B=16, x{TRAINS_PER_STEP} (fp32, '
                    f'{threads} torch threads)', 'ms_per_step': t * 1e3, 'parts': parts}


def run_reference(args):
  if int(os.environ.get('RANK', 0)) != 0:
    return
  steps = max(1, min(args.steps, 3))
  warm = min(args.warmup, 1)
  learner = None if args.agent == 'feed' else OracleLearner(args.size)
  for _ in range(warm):
    if learner is not None:
      learner.time_policy(8)
  times = [cpu_iteration(args, learner) for _ in range(steps)]
  t = float(np.mean([x[0] for x in times]))
  v = NENVS / t
  threads = times[0][2]
  print(json.dumps({
      'impl': 'reference', 'metric': 'env_steps_per_sec', 'value': v, 'unit': 'env steps/s',
      'learner_samples_per_sec': v * TRAIN_RATIO,
      'n_gpus': args.gpus, 'steps': steps, 'warmup': warm,
      'ms_per_step': t * 1e3, 'higher_is_better': True, 'scaling': 'weak',
      'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
      'config': {'workload': workload_name(args) + ' -- CPU restatement of the reference '
                             '(oracle/), see cpu_baseline.sample', 'envs': NENVS, 'batch': [B, T]},
      'cpu_baseline': {'value': v, 'unit': 'env steps/s', 'cores': threads, 'kind': 'port',
                       'sample': cpu_iteration.__doc__.strip().replace('\n  ', ' '),
                       'parts': times[-1][1]},
      'e2e': {'value': v, 'unit': 'env steps/s', 'h2d_bytes_per_step': 0,
              'd2h_bytes_per_step': 0}}), flush=True)


def main():
  if os.environ.get('EMB_BENCH_DEBUG'):       # hang diagnosis: dump all stacks periodically
    import faulthandler
    faulthandler.dump_traceback_later(int(os.environ['EMB_BENCH_DEBUG']), repeat=True, file=sys.stderr)
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=10)
  ap.add_argument('--warmup', type=int, default=3)
  ap.add_argument('--agent', default='dreamerv3', choices=['dreamerv3', 'feed'])
  ap.add_argument('--size', default='size200m')
  ap.add_argument('--dtype', default='bfloat16', choices=['bfloat16', 'float32'])
  ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
  ap.add_argument('--capacity', type=float, default=2e5)
  ap.add_argument('--no-cpu', action='store_true')
  args = ap.parse_args()
  args.warmup = max(args.warmup, 3)
  if args.impl == 'reference':
    run_reference(args)
  else:
    run_b200(args)


if __name__ == '__main__':
  main()
