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
NCU_TRAFFIC = {'rssm_fwd': 8866091000 + 192772352,      # profiles/r02_rssm_tma_kernels_ncu.md
               'rssm_bwd': 10102213000 + 138353664}


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
                                graph=os.environ.get('EMB_GRAPH', 'auto'),
                                grad_buckets={'off': False, 'on': True}.get(
                                    os.environ.get('EMB_GRAD_BUCKETS', 'auto'), 'auto')))
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


def _fallbacks(args):
  if args.agent == 'feed':
    return {}
  from embodied_b200.dreamerv3 import ops
  return dict(ops.FALLBACKS)


def conv_lines(torch, cfg, flush):
  """The tcgen05 convolution kernels (csrc/conv_tc.cu) at this model's layer shapes, B*T = 1024
  images: each launch timed alone with CUDA events after the timed region (the captured train
  step replays ~50 of them back to back; a stopwatch node around each would perturb the graph).
  Tensor-bound: achieved = 2 * pixels * taps * Cin * Cout / time against the measured bf16 peak."""
  from embodied_b200.dreamerv3 import ops
  try:
    peak = float(json.load(open(ROOT / 'MEASURED_PEAKS.json'))['bf16_tflops'])
    src = 'measured burst (MEASURED_PEAKS.json)'
  except Exception:
    peak, src = 1590.0, 'fallback (B200_PROFILING.md)'
  depths = [cfg.depth * m for m in cfg.mults]
  n = B * T
  res = cfg.image[0]
  layers = []
  for i in range(1, len(depths)):                     # encoder: conv on the pooled grid of stage i - 1
    layers.append((f'enc/cnn{i} fwd', res >> i, depths[i - 1], depths[i], 5))
  out = []
  for name, hw, cin, cout, k in layers:
    x = torch.randn((n, hw, hw, cin), device='cuda').to(torch.bfloat16)
    if not ops.conv_tc_supported(x, cin, cout, k):
      continue
    wp = ops.pack_conv_weight((torch.randn((k, k, cin, cout), device='cuda') / 50).to(torch.bfloat16))
    for _ in range(3):
      ops.conv_tc(x, wp, k=k)
    ts = []
    for _ in range(5):
      flush.zero_()
      a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      a.record(); ops.conv_tc(x, wp, k=k); b.record()
      torch.cuda.synchronize()
      ts.append(a.elapsed_time(b))
    t = float(np.median(ts)) * 1e-3
    flops = 2.0 * n * hw * hw * k * k * cin * cout
    out.append({'kernel': f'conv_tc_kernel ({name}, {hw}x{hw}, {cin}->{cout}, {n} images)',
                'bound': 'tensor', 'achieved': flops / t / 1e12, 'peak': peak, 'peak_source': src,
                'unit': 'TFLOP/s', 'frac': flops / t / 1e12 / peak, 'us_per_launch': t * 1e6,
                'launches_timed': 5, 'timed': 'alone, after the timed region', 'traffic': None})
  return out


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
  if args.scaling == 'strong':
    # the GLOBAL workload is fixed -- 256 envs and a (16, 64) batch in all; rank r owns
    # envs {i : i mod R = r} and samples 16 / R windows from its own shard per update
    global B, NENVS
    assert 16 % world == 0, f'strong scaling splits B=16 over the ranks; {world} does not divide it'
    B, NENVS = 16 // world, 256 // world
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
  if args.agent != 'feed' and args.dtype == 'bfloat16':
    kernels += conv_lines(torch, loop.agent.cfg, flush)
  kernels = [k for k in kernels if k]
  # the roofline line is the hand-written kernel with the largest share of the step
  per_step = {'rows_kernel': TRAINS_PER_STEP, 'rssm_bwd_kernel': TRAINS_PER_STEP,
              'rssm_fwd_kernel': TRAINS_PER_STEP}
  share = lambda k: k['us_per_launch'] * per_step.get(k['kernel'].split()[0], 0) * args.steps
  roofline = max(kernels, key=share)
  for k in kernels:
    if k['kernel'].split()[0] in per_step:
      k['share_of_step'] = share(k) * 1e-6 / t_dev

  exchange = None
  ex = getattr(loop.agent, 'exchange', None)
  if ex is not None:
    tail = [t for k, t in prof if k == 'exchange_tail']
    exchange = {
        'buckets': [{'group': b['group'], 'MB': round(b['elem_count'] * 4 / 1e6, 1)} for b in ex.buckets],
        'payload_MB_per_update': round(sum(b['elem_count'] for b in ex.buckets) * 4 / 1e6, 1),
        'exposed_ms_per_update': (float(np.mean([max(t, 0.0) for t in tail])) * 1e-3 if tail else None),
        'note': 'all-reduce (avg, own NCCL communicator) + optimiser per bucket on a side stream under '
                'the backward pass; exposed = side-stream end minus backward end, CUDA events inside '
                'the captured graph, rank 0'}
  env_steps = world * NENVS * args.steps
  h2d = NENVS * (int(np.prod(IMAGE)) + 4 + 3 + 20 + 8) + TRAINS_PER_STEP * (B * L * 8 + B * T * 8)
  d2h = NENVS * 4 + 4
  line = {
      'metric': 'env_steps_per_sec', 'value': env_steps / t_dev, 'unit': 'env steps/s',
      'learner_samples_per_sec': env_steps / t_dev * TRAIN_RATIO,
      'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
      'ms_per_step': t_dev / args.steps * 1e3, 'higher_is_better': True,
      'scaling': args.scaling, 'vs_baseline': None,
      'dtype': 'u8' if args.agent == 'feed' else {'bfloat16': 'bf16', 'float32': 'f32'}[args.dtype],
      'data': 'synthetic',
      'config': {
          'workload': workload_name(args),
          'envs_per_gpu': NENVS, 'batch': [B, T], 'train_ratio': TRAIN_RATIO,
          'learner_steps_per_step': TRAINS_PER_STEP, 'replay_capacity_items': capacity,
          'parallelism': f'dp{world}: envs, replay shard and batch per rank; NCCL grad all-reduce',
          'precision': 'bf16 compute with fp32 master weights, norms and losses (the reference default, '
                       'embodied/jax/nets.py:12); the CPU arm (--impl reference, cpu_baseline) computes in fp32',
          'cache': 'L2 flushed (256 MiB write) before every timed step; replay tables > L2'},
      'e2e': {'value': env_steps / t_e2e, 'unit': 'env steps/s',
              'ms_per_step': t_e2e / args.steps * 1e3,
              'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h},
      'gpu_launches': launches,
      # call sites of the learner that ran a library formulation because an own kernel does not
      # take their shape (embodied_b200/dreamerv3/model.py Model._use); {} = none
      'eager_fallbacks': _fallbacks(args),
      'roofline': roofline, 'kernels': kernels, 'exchange': exchange,
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


# ------------------------------------------- config 5: replay sample-rate sweep
SWEEP_B = (8, 16, 32, 64, 128)
SWEEP_T = (16, 32, 64, 128, 256)
ROW_KINDS = {                      # bytes per stored step, excluding the 20-byte stepid
    'image': {'image': (np.uint8, IMAGE), 'is_first': (bool, ())},     # Consec reads is_first
    'default': {'image': (np.uint8, IMAGE), 'dyn/deter': (np.float32, (DETER,)),
                'dyn/stoch': (np.float32, STOCH), 'reward': (np.float32, ()),
                'action': (np.int32, ()), 'is_first': (bool, ()), 'is_last': (bool, ()),
                'is_terminal': (bool, ())},
}


def run_sweep(args):
  """BASELINE.json config 5: Replay.sample through Stateless -> Consec for
  B x T windows, image-only rows (12 288 B) and default dreamerv3 rows (53 279 B), buffer
  pre-filled with 8*B*T steps from 64 workers; one independent shard per rank.  A step is ONE
  sampled batch: samples/s = B*T / step, GB/s = 2*B*(T+1)*(row bytes + 24) / step."""
  import torch
  import torch.distributed as dist
  import embodied_b200 as embodied
  from embodied_b200 import _lib
  rank, world, local = (int(os.environ.get(k, d)) for k, d in (('RANK', 0), ('WORLD_SIZE', 1), ('LOCAL_RANK', 0)))
  torch.cuda.set_device(local)
  if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
  peak, peak_src = peaks()
  flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
  g = torch.Generator(device='cuda').manual_seed(rank)
  workers = 64
  points, launches0 = [], _lib.launch_count()
  grid = [(b, t) for b in SWEEP_B for t in SWEEP_T]
  if args.sweep_points:
    grid = [tuple(int(x) for x in p.split('x')) for p in args.sweep_points.split(',')]
  with ClockSampler(local) as clocks:
    for kind, spec in ROW_KINDS.items():
      rowbytes = sum(int(np.dtype(d).itemsize * np.prod(sh, dtype=np.int64)) for d, sh in spec.values())
      for Bx, Tx in grid:
        Lx = Tx + PREFIX
        per_worker = max(8 * Bx * Tx // workers, 2 * Lx)
        replay = embodied.Replay(Lx, None, chunksize=1024, seed=0, staging_rows=workers, workers=workers)
        torch_dt = {np.uint8: torch.uint8, np.float32: torch.float32, np.int32: torch.int32, bool: torch.bool}
        step = {}
        for k, (d, sh) in spec.items():
          if d is bool:
            step[k] = torch.zeros((workers, *sh), dtype=torch.bool, device='cuda')
          elif d is np.uint8:
            step[k] = torch.randint(0, 256, (workers, *sh), generator=g, device='cuda', dtype=torch.uint8)
          elif d is np.int32:
            step[k] = torch.randint(0, CLASSES, (workers, *sh), generator=g, device='cuda', dtype=torch.int32)
          else:
            step[k] = torch.randn((workers, *sh), generator=g, device='cuda')
        for _ in range(per_worker):
          replay.add_batch(step)
        base = embodied.streams.Stateless(replay.sample, Bx, 'train')
        stream = iter(embodied.streams.Consec(base, length=Tx, consec=1, prefix=PREFIX, strict=True,
                                              contiguous=True))
        for _ in range(max(args.warmup, 3)):
          batch = next(stream)
        torch.cuda.synchronize()
        if world > 1:
          dist.barrier()
        total = 0.0
        for _ in range(args.steps):
          flush.zero_()
          a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
          torch.cuda.synchronize()
          a.record()
          batch = next(stream)
          b.record()
          torch.cuda.synchronize()
          total += a.elapsed_time(b) * 1e-3
        t = torch.tensor([total], device='cuda', dtype=torch.float64)
        if world > 1:
          dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item()) / args.steps
        nbytes = 2 * Bx * Lx * (rowbytes + 20 + 4)
        points.append({'rows': kind, 'row_bytes': rowbytes, 'B': Bx, 'T': Tx,
                       'us_per_batch': dt * 1e6, 'samples_per_sec': world * Bx * Tx / dt,
                       'GBs_per_gpu': nbytes / dt / 1e9, 'frac': nbytes / dt / 1e9 / peak})
        del replay, stream, base, batch
        torch.cuda.empty_cache()
  head = next((p for p in points if p['rows'] == 'default' and (p['B'], p['T']) == (B, T)), points[-1])
  best = max(points, key=lambda p: p['frac'])
  line = {
      'metric': 'replay_samples_per_sec', 'value': head['samples_per_sec'], 'unit': 'samples/s',
      'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
      'ms_per_step': head['us_per_batch'] * 1e-3, 'higher_is_better': True, 'scaling': 'weak',
      'vs_baseline': None, 'dtype': 'u8', 'data': 'synthetic',
      'config': {'workload': 'config 5: replay sample-rate sweep B in {8..128} x T in {16..256}, 64x64x3 '
                             'obs, image-only and default dreamerv3 rows; headline = default rows at '
                             f'B={head["B"]}, T={head["T"]}',
                 'parallelism': f'{world} independent replay shards, no collective',
                 'cache': 'L2 flushed (256 MiB write) before every timed batch; host index work + '
                          'row-id H2D + gather launch inside the timed region'},
      'roofline': {'kernel': 'rows_kernel (emb_replay_gather)', 'bound': 'hbm',
                   'achieved': head['GBs_per_gpu'], 'peak': peak, 'peak_source': peak_src, 'unit': 'GB/s',
                   'frac': head['frac'], 'traffic': None,
                   'note': 'whole Replay.sample call (host draw + launch), not the bare kernel'},
      'best_point': best, 'sweep': points,
      'gpu_launches': _lib.launch_count() - launches0, 'clocks': clocks.summary(),
      'e2e': {'value': head['samples_per_sec'], 'unit': 'samples/s',
              'h2d_bytes_per_step': head['B'] * (head['T'] + PREFIX) * 8, 'd2h_bytes_per_step': 0,
              'note': 'the sampled batch stays in HBM by design (the learner consumes it there); '
                      'H2D = the int64 row ids'}}
  if rank == 0:
    print(json.dumps(line), flush=True)
  if world > 1:
    dist.barrier()
    dist.destroy_process_group()


# ---------------------------------------------- configs 3 / 4: the Driver + Replay half
ROWS_WORKLOADS = {
    # name: (BASELINE config, envs in all, shards the config names, env factory, policy kind)
    'atari_rows': ('config 3: Atari-shape image u8[84,84,1], 18 discrete actions, 1024 envs, replay sharded '
                   'by worker (Driver + Replay half; the dreamerv3 update of this config runs at 96x96, SURVEY F7)',
                   1024, 8),
    'proprio_rows': ('config 4: DMC-proprio-shape dummy (orientations f32[14], height f32[], velocity f32[9], '
                     'action f32[6]), 512 envs, Driver + Replay + a host old-API policy (the director agent\'s '
                     'role; its networks are not built)', 512, 4),
}


def rows_cpu_line(workload, seconds=8.0):
  """The reference's CPU path for a rows workload on this host: oracle Driver (serial, the
  reference default) + OracleReplay.add per transition + sample(B) -> Consec view every
  B*T/train_ratio env steps, same synthetic envs, a host random policy; timed over >= `seconds`
  of steady state with time.perf_counter.  Single-threaded Python like the reference."""
  import itertools
  from oracle import host_oracle as ho
  from embodied_b200.envs import synthetic
  name, total_envs, shards = ROWS_WORKLOADS[workload]
  atari = workload == 'atari_rows'
  n = total_envs
  if atari:
    envs = [synthetic.SyntheticImage(i, size=(84, 84, 1), classes=18, length=500) for i in range(n)]
  else:
    envs = [synthetic.SyntheticProprio(i, length=500) for i in range(n)]
  replay = ho.OracleReplay(L, int(2e5), 1024, True, 0, ids=itertools.count(1))
  driver = ho.OracleDriver(envs, {k: v for k, v in envs[0].act_space.items() if k != 'reset'})
  driver.callbacks.append(lambda row, w: replay.add(row, w))
  rng = np.random.default_rng(0)
  acts = (rng.integers(0, 18, (4, n)).astype(np.int32) if atari else
          rng.uniform(-1, 1, (4, n, 6)).astype(np.float32))
  state = {'t': 0}

  def policy(carry, obs):
    if atari:
      ho.normalize_image(obs['image'])                       # the agent's input cast (rssm.py:230)
    state['t'] += 1
    return carry, {'action': acts[state['t'] % 4]}, {}
  per_step = max(1, TRAIN_RATIO * n // (B * T))
  while len(replay) < 4 * B * L:
    driver.step(policy)
  t0, steps = time.perf_counter(), 0
  while time.perf_counter() - t0 < seconds:
    driver.step(policy)
    for _ in range(per_step):
      ho.consec_view(replay.sample(B), T, 0, PREFIX)
    steps += 1
  dt = (time.perf_counter() - t0) / steps
  return {'value': n / dt, 'unit': 'env steps/s', 'cores': 1, 'kind': 'port',
          'sample': f'{steps} Driver steps over {n} envs + {per_step} x (Replay.sample({B}) + Consec view) each, '
                    f'{seconds:.0f} s of steady state; oracle/host_oracle.py = numpy restatement of '
                    'embodied/core/{driver,replay,chunk,selectors,streams}.py (pinned byte for byte against '
                    'the reference files, tests/test_oracle_pinned.py)', 'seconds_per_step': dt}


def run_rows_reference(args):
  line = rows_cpu_line(args.workload, seconds=max(5.0, 2.0 * (args.steps + args.warmup)))
  print(json.dumps({
      'impl': 'reference', 'metric': 'env_steps_per_sec', 'value': line['value'], 'unit': 'env steps/s',
      'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
      'ms_per_step': line['seconds_per_step'] * 1e3, 'higher_is_better': True, 'scaling': 'strong',
      'vs_baseline': None, 'dtype': 'u8', 'data': 'synthetic',
      'config': {'workload': ROWS_WORKLOADS[args.workload][0] + ' -- CPU restatement of the reference (oracle/)'},
      'cpu_baseline': line,
      'e2e': {'value': line['value'], 'unit': 'env steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}),
      flush=True)


def run_rows(args):
  """BASELINE configs 3 and 4, the part of them this library computes: N envs stepped on the
  host -> pinned staging -> ONE H2D -> emb_driver_stage_obs (append + normalise) -> a device
  policy stand-in -> emb_driver_scatter_mask_actions, and one Replay.sample(B=16) + Consec per
  `train_ratio` env steps (emb_replay_gather).  One step = one Driver step over this rank's envs.
  `value`: the step with observations resident in HBM; `e2e`: through Driver(...) over host envs."""
  import torch
  import torch.distributed as dist
  import embodied_b200 as embodied
  from embodied_b200 import _lib, elements
  from embodied_b200.envs import synthetic
  rank, world, local = (int(os.environ.get(k, d)) for k, d in (('RANK', 0), ('WORLD_SIZE', 1), ('LOCAL_RANK', 0)))
  torch.cuda.set_device(local)
  if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
  peak, peak_src = peaks()
  name, total_envs, shards = ROWS_WORKLOADS[args.workload]
  n = total_envs // world                                       # strong split: the config fixes the env count
  atari = args.workload == 'atari_rows'
  if atari:
    mk = lambda i: synthetic.SyntheticImage(rank * n + i, size=(84, 84, 1), classes=18, length=500)
  else:
    mk = lambda i: synthetic.SyntheticProprio(rank * n + i, length=500)
  env = mk(0)
  obs_space, act_space = env.obs_space, env.act_space
  g = torch.Generator(device='cuda').manual_seed(rank)

  class Policy:                                                  # device policy stand-in, Agent protocol
    device_obs = True
    ext_space = {}
    def __init__(self):
      if atari:
        self.acts = torch.randint(0, 18, (4, n), generator=g, device='cuda', dtype=torch.int32)
      else:
        self.acts = torch.rand((4, n, 6), generator=g, device='cuda') * 2 - 1
      self.t = 0
    def init_policy(self, k):
      return ()
    def policy(self, carry, obs, mode='train'):
      self.t += 1
      return carry, {'action': self.acts[self.t % 4]}, {}

  agent = Policy()
  replay = embodied.Replay(L, int(args.capacity), chunksize=1024, online=True, seed=0,
                           staging_rows=n, workers=n)
  stream = iter(embodied.streams.Consec(embodied.streams.Stateless(replay.sample, B, 'train'),
                                        length=T, consec=1, prefix=PREFIX, strict=True, contiguous=True))
  driver = embodied.Driver([(lambda i=i: mk(i)) for i in range(n)], parallel=False, fetch_outs=False)
  driver.on_step(replay.add)
  samples_per_step = max(1, TRAIN_RATIO * n // (B * T))
  state = {'on': False, 'last': None}

  def sample(trans=None, k=None):
    if state['on']:
      for _ in range(samples_per_step):
        state['last'] = next(stream)
  driver.on_batch(sample)
  driver.reset(agent.init_policy)
  while len(replay) < 4 * B * L:
    driver(agent.policy, steps=n)
  state['on'] = True
  rowbytes = replay.store.bytes_per_row
  res = {k: (torch.randint(0, 256, (n, *s.shape), generator=g, device='cuda', dtype=torch.uint8)
             if s.dtype == np.uint8 else
             torch.zeros((n, *s.shape), dtype=torch.bool, device='cuda') if s.dtype == bool else
             torch.randn((n, *s.shape), generator=g, device='cuda'))
         for k, s in obs_space.items()}
  flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')

  def step_e2e():
    driver(agent.policy, steps=n)
    return state['last']['reward'][0, 0].cpu()

  def step_resident():
    _, acts, _ = agent.policy((), res)
    replay.add_batch({**res, **acts})
    sample()

  def timed(fn):
    for _ in range(max(args.warmup, 3)):
      fn()
    torch.cuda.synchronize()
    if world > 1:
      dist.barrier()
    total = 0.0
    for _ in range(args.steps):
      flush.zero_()
      torch.cuda.synchronize()
      t0 = time.perf_counter()
      a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      a.record()
      fn()
      b.record()
      torch.cuda.synchronize()
      # the step is host-driven (env stepping, index bookkeeping): wall time of the synchronised
      # step, never less than the device span
      total += max(time.perf_counter() - t0, a.elapsed_time(b) * 1e-3)
    t = torch.tensor([total], device='cuda', dtype=torch.float64)
    if world > 1:
      dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()) / args.steps

  launches0 = _lib.launch_count()
  with ClockSampler(local) as clocks:
    dt = timed(step_resident)
    launches = _lib.launch_count() - launches0
    dt_e2e = timed(step_e2e)
  # the gather launch alone, CUDA events on its stream, L2 flushed
  us = []
  for _ in range(10):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    batch = next(stream)
    b.record()
    torch.cuda.synchronize()
    us.append(a.elapsed_time(b) * 1e3)
  gather_bytes = 2 * B * L * (rowbytes + 4)
  med = sorted(us)[len(us) // 2]
  h2d = sum(int(np.dtype(s.dtype).itemsize * np.prod(s.shape, dtype=np.int64)) for s in obs_space.values()) * n
  line = {
      'metric': 'env_steps_per_sec', 'value': world * n / dt, 'unit': 'env steps/s', 'n_gpus': world,
      'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': dt * 1e3,
      'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'u8', 'data': 'synthetic',
      'config': {'workload': name, 'envs_per_gpu': n, 'batch': [B, T], 'row_bytes': rowbytes,
                 'samples_per_step': samples_per_step,
                 'parallelism': f'{world} replay shards by worker, no collective',
                 'cache': 'L2 flushed (256 MiB write) before every timed step'},
      'e2e': {'value': world * n / dt_e2e, 'unit': 'env steps/s', 'ms_per_step': dt_e2e * 1e3,
              'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': 4 + n * (4 if atari else 24)},
      'gpu_launches': launches,
      'roofline': {'kernel': f'rows_kernel (emb_replay_gather, B={B} L={L}, whole Replay.sample call)',
                   'bound': 'hbm', 'achieved': gather_bytes / (med * 1e-6) / 1e9, 'peak': peak,
                   'peak_source': peak_src, 'unit': 'GB/s', 'frac': gather_bytes / (med * 1e-6) / 1e9 / peak,
                   'us_per_launch': med, 'algorithmic_bytes': gather_bytes, 'traffic': None},
      'clocks': clocks.summary(),
      'cpu_baseline': None}
  if rank == 0:
    if world == 1 and not args.no_cpu:
      line['cpu_baseline'] = rows_cpu_line(args.workload)
    print(json.dumps(line), flush=True)
  if world > 1:
    dist.barrier()
    dist.destroy_process_group()


def workload_name(args):
  if args.agent == 'feed':
    return ('config 2 PLUMBING ONLY: Driver(256 envs, 64x64x3 u8) + Replay(L=65, 53 299 B rows '
            'incl. f32 latents) append/sample(B=16)/update; no model, latents from a device pool')
  if IMAGE != (64, 64, 3):
    return (f'config 3 (learner half): dreamerv3 {args.size} on a synthetic {"x".join(map(str, IMAGE))} image env '
            f'({CLASSES} actions), 256 envs per GPU, replay (B=16,T=64,L=65), one Driver step + 8 x (sample -> '
            'train -> update)')
  return (f'config 2: dreamerv3 {args.size} on synthetic 64x64x3 image env, 256 envs, '
          'replay (B=16,T=64,L=65), one Driver step + 8 x (sample -> train -> update)')


# ------------------------------------------------------ the reference (CPU) arm
# One benchmark step = 256 env steps + 8 learner updates of (B=16, T=64): ~2 minutes of host time.
# A SAMPLE is 1/8 of it measured whole, nothing extrapolated: one learner update on a real
# (B=16, T) batch together with the B*T/32 env steps that trigger it (train_ratio = 32:
# embodied/run/train.py:25-26), i.e. for T = 64: 32 env steps of Driver + Replay + policy, one
# replay.sample(16) + Consec view + replay.update, one fp32 dreamerv3 update of 1024 samples.
# value = env steps in the sample / its wall time.  When more steps are requested than fit the
# time budget, later samples shorten T (64 -> 32 -> 16 -> 8; every per-sample cost is linear in T
# except the optimiser pass over the parameters, which makes short samples slightly PESSIMISTIC
# for the CPU, so the reported figure is the T = 64 sample whenever one was timed).
class OracleSample:
  """The host-side restatement (oracle/) of one sample.  bench.py is one of the few places
  allowed to execute oracle/; it is the thing compared against, never shipped."""

  def __init__(self, size, T_s, rank=0, agent='dreamerv3'):
    import itertools
    from oracle import host_oracle as ho
    self.ho, self.T = ho, T_s
    self.L = T_s + PREFIX
    self.nenvs = max(1, B * T_s // TRAIN_RATIO)
    envs = [make_env(rank * NENVS + i) for i in range(self.nenvs)]
    self.replay = ho.OracleReplay(self.L, None, 1024, True, 0, ids=itertools.count(1))
    self.driver = ho.OracleDriver(envs, envs[0].act_space)
    self.driver.callbacks.append(lambda row, w: self.replay.add(row, w))
    self.learner = None if agent == 'feed' else _learner(size)
    rng = np.random.default_rng(rank)
    deter, stoch = DETER, STOCH
    if self.learner is not None:
      deter, stoch = self.learner.cfg.deter, (self.learner.cfg.stoch, self.learner.cfg.classes)
    self.deter = rng.standard_normal((self.nenvs, deter), dtype=np.float32)
    self.stoch = rng.standard_normal((self.nenvs, *stoch), dtype=np.float32)
    self.action = rng.integers(0, CLASSES, self.nenvs).astype(np.int32)
    self.pcarry = None
    self.learn = False
    while len(self.replay) < 4 * B:              # prefill: plumbing only
      self.driver.step(self._feed_policy)
    self.learn = True

  def _feed_policy(self, carry, obs):
    self.ho.normalize_image(obs['image'])
    return carry, {'action': self.action}, {'dyn/deter': self.deter, 'dyn/stoch': self.stoch}

  def _policy(self, carry, obs):
    """Reference Agent.policy on the host: encoder + one RSSM step + actor (fp32 oracle)."""
    torch, cfg, n = self.learner.torch, self.learner.cfg, self.nenvs
    if self.pcarry is None:
      self.pcarry = dict(deter=torch.zeros(n, cfg.deter), stoch=torch.zeros(n, cfg.stoch, cfg.classes),
                         action=torch.zeros(n, dtype=torch.int32))
    noise = dict(stoch=torch.zeros(n, cfg.stoch, cfg.classes), action=torch.zeros(n, cfg.actions))
    self.pcarry, act, out = self.learner.model.policy(
        self.pcarry, torch.from_numpy(obs['image']), torch.from_numpy(np.asarray(obs['is_first'])), noise)
    return carry, {'action': act['action'].numpy()}, {k: v.numpy() for k, v in out.items()}

  def step(self):
    t0 = time.perf_counter()
    self.driver.step(self._feed_policy if self.learner is None else self._policy)
    batch = self.ho.consec_view(self.replay.sample(B), self.T, 0, PREFIX)
    if self.learner is not None:
      self.learner.train_on(batch, self.T)
    self.replay.update({k: batch[k][:, PREFIX:] for k in ('stepid', 'dyn/deter', 'dyn/stoch')})
    return time.perf_counter() - t0


_LEARNERS = {}


def _learner(size):
  if size not in _LEARNERS:
    _LEARNERS[size] = OracleLearner(size)
  return _LEARNERS[size]


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
    self.cfg = do.default_config(**dict(C.SIZES[size], image=tuple(IMAGE), actions=CLASSES))
    self.model = do.Dreamer(self.cfg, do.init_params(self.cfg, 0))
    self.noise = {}

  def train_on(self, batch, T_s):
    """One full update (forward, backward, optimiser) on a replay batch of numpy arrays."""
    torch = self.torch
    data = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in batch.items()}
    if T_s not in self.noise:
      self.noise[T_s] = self.do.make_noise(self.cfg, B, T_s, 0)
    self.model.train(data, self.noise[T_s])


def cpu_sample_line(size, agent, T_s, seconds, threads):
  nenv = max(1, B * T_s // TRAIN_RATIO)
  return {
      'value': nenv / seconds, 'unit': 'env steps/s', 'cores': threads, 'kind': 'port',
      'sample': (f'1/{TRAINS_PER_STEP * T // T_s} of one benchmark step, measured whole: {nenv} env steps of the '
                 f'oracle Driver + Replay + fp32 dreamerv3 policy, replay.sample(16) + Consec + replay.update, '
                 f'and ONE fp32 dreamerv3 update on the sampled (B=16, T={T_s}) batch ({size}, '
                 f'{threads} torch threads; the Driver / Replay half is single-threaded Python like the '
                 f'reference).  oracle/ = restatement of the reference (JAX is not installable here)'
                 if agent != 'feed' else
                 f'{nenv} env steps of the oracle Driver + Replay, replay.sample(16) + Consec + update'),
      'seconds_per_sample': seconds}


def cpu_baseline(args):
  """~15-30 s of host work: one T = 64 sample after a T = 8 warm-up sample."""
  if args.agent != 'feed':
    OracleSample(args.size, 8, agent=args.agent).step()         # thread pools, allocator warm-up
  sample = OracleSample(args.size, T, agent=args.agent)
  dt = sample.step()
  threads = sample.learner.threads if sample.learner else 1
  return cpu_sample_line(args.size, args.agent, T, dt, threads)


def run_reference(args):
  if int(os.environ.get('RANK', 0)) != 0:
    return
  budget = float(os.environ.get('EMB_REF_BUDGET_S', 150))
  total = args.steps + args.warmup
  samples = {T: OracleSample(args.size, T, agent=args.agent)}
  t_first = samples[T].step()                       # warm-up step 1 doubles as the calibration
  # T of the remaining samples: the longest that fits the budget (cost ~ 10 % fixed + 90 % linear in T)
  T_s = T
  while T_s > 8 and (total - 1) * t_first * (0.1 + 0.9 * T_s / T) > budget:
    T_s //= 2
  if T_s not in samples:
    samples[T_s] = OracleSample(args.size, T_s, agent=args.agent)
  for _ in range(args.warmup - 1):
    samples[T_s].step()
  times = [samples[T_s].step() for _ in range(args.steps)]
  t = float(np.mean(times))
  nenv = samples[T_s].nenvs
  threads = samples[T].learner.threads if samples[T].learner else 1
  line = cpu_sample_line(args.size, args.agent, T_s, t, threads)
  v = nenv / t
  print(json.dumps({
      'impl': 'reference', 'metric': 'env_steps_per_sec', 'value': v, 'unit': 'env steps/s',
      'learner_samples_per_sec': v * TRAIN_RATIO,
      'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
      'ms_per_step': t * 1e3, 'higher_is_better': True, 'scaling': 'weak',
      'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
      'config': {'workload': workload_name(args) + ' -- CPU restatement of the reference (oracle/); a '
                             'reference-arm step is a bounded sample of the benchmark step, see '
                             'cpu_baseline.sample', 'envs': NENVS, 'batch': [B, T],
                 'sample_T': T_s, 'env_steps_per_sample': nenv},
      'cpu_baseline': dict(line, value=v, first_sample_T64_seconds=t_first,
                           first_sample_T64_value=samples[T].nenvs / t_first),
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
  ap.add_argument('--scaling', default='weak', choices=['weak', 'strong'],
                  help='weak: 256 envs and a (16, 64) batch PER GPU; strong: in all (SURVEY 8e)')
  ap.add_argument('--workload', default='train', choices=['train', 'replay_sweep', 'atari_rows', 'proprio_rows'],
                  help='train = BASELINE config 2 (the headline); replay_sweep = config 5; atari_rows / '
                       'proprio_rows = the Driver + Replay half of configs 3 / 4')
  ap.add_argument('--image', default='', help='HxWxC of the synthetic observation for the train workload '
                  '(default 64x64x3 = config 2; 96x96x1 = the dreamerv3 update of config 3, SURVEY F7)')
  ap.add_argument('--classes', type=int, default=0, help='discrete actions of the synthetic env (default 5; Atari: 18)')
  ap.add_argument('--sweep-points', default='', help='e.g. 16x64,128x256 (default: the full 5x5 grid)')
  args = ap.parse_args()
  args.warmup = max(args.warmup, 3 if args.impl != 'reference' else 1)
  global IMAGE, CLASSES, ROW_BYTES
  if args.image:
    IMAGE = tuple(int(x) for x in args.image.lower().split('x'))
    ROW_BYTES += int(np.prod(IMAGE)) - 12288
  if args.classes:
    CLASSES = args.classes
  if args.impl == 'reference':
    if int(os.environ.get('RANK', 0)) != 0:
      return                                   # rank 0 alone runs the CPU arm
    run_rows_reference(args) if args.workload in ROWS_WORKLOADS else run_reference(args)
  elif args.workload == 'replay_sweep':
    run_sweep(args)
  elif args.workload in ROWS_WORKLOADS:
    run_rows(args)
  else:
    run_b200(args)


if __name__ == '__main__':
  main()
