"""DreamerV3 agent behind the embodied Agent protocol (embodied/core/base.py:1-31).

Same constructor and methods as the reference's ``dreamerv3.Agent`` wrapped by
``embodied.jax.Agent`` (dreamerv3/agent.py:24-154, embodied/jax/agent.py:220-337):
``Agent(obs_space, act_space, config)``, ``init_policy/init_train/init_report``,
``policy(carry, obs, mode) -> (carry, act, out)``, ``train(carry, data) ->
(carry, outs, metrics)`` with ``outs['replay'] = {stepid, dyn/deter, dyn/stoch}``,
``report``, ``stream``, ``save``, ``load``, ``ext_space``, ``policy_keys``.

Differences, all about where bytes live: observations arrive as device tensors
staged by ``emb_driver_stage_obs`` (``device_obs = True``; uint8 images also
normalised to float32 there), actions / latents are returned as device tensors
and appended to the replay by ``emb_driver_scatter_mask_actions`` without
visiting the host, train batches are the dense device tensors of
``emb_replay_gather`` and ``outs['replay']`` stays on the device for
``emb_replay_scatter_update``.  With ``torch.distributed`` initialised, the flat
gradient buffer is averaged with ONE NCCL all-reduce (embodied/jax/opt.py:52-54).
"""
import numpy as np
import torch

from .. import elements
from ..core import base
from . import config as configlib
from . import model as modellib
from . import optim
from . import params as paramlib

f32 = torch.float32


def _clone(tree):
  if isinstance(tree, dict):
    return {k: _clone(v) for k, v in tree.items()}
  if isinstance(tree, (tuple, list)):
    return type(tree)(_clone(v) for v in tree)
  return tree.clone()


def _copy_into(dst, src):
  """dst <- src leaf by leaf; a bare tensor stands for a single-key dict."""
  if isinstance(dst, dict):
    if not isinstance(src, dict):
      assert len(dst) == 1, list(dst)
      src = {next(iter(dst)): src}
    for k in dst:
      _copy_into(dst[k], src[k])
  elif isinstance(dst, (tuple, list)):
    for a, b in zip(dst, src):
      _copy_into(a, b)
  else:
    dst.copy_(src)


def _detach(tree):
  if isinstance(tree, dict):
    return {k: _detach(v) for k, v in tree.items()}
  return tree.detach() if isinstance(tree, torch.Tensor) else tree


class Agent(base.Agent):

  device_obs = True

  def __init__(self, obs_space, act_space, config=None, device=None, values=None):
    if not torch.cuda.is_available():
      raise RuntimeError(
          'embodied_b200.dreamerv3.Agent runs on a CUDA device; none is visible '
          'and there is no CPU fallback.')
    self.obs_space = dict(obs_space)
    self.act_space = {k: v for k, v in act_space.items() if k != 'reset'}
    cfg = config if isinstance(config, configlib.Config) else configlib.make(**(config or {}))
    from . import spaces as spacelib
    cfg = configlib.Config(cfg)
    info = spacelib.analyze(self.obs_space, self.act_space)
    if not info['actspec']:
      raise ValueError('the action space has no key besides `reset`')
    cfg.update(image=info['image'], imgkeys=info['imgkeys'], vecspec=info['vecspec'],
               actspec=info['actspec'], actions=sum(spacelib.width(a) for a in info['actspec']))
    self.obskeys = [k for k, _ in info['imgkeys']] + [v[0] for v in info['vecspec']]
    self.actkeys = [a[0] for a in info['actspec']]
    self.cfg = cfg
    self.device = torch.device(device if device is not None else
                               f'cuda:{torch.cuda.current_device()}')
    self.cd = {'bfloat16': torch.bfloat16, 'float32': f32}[cfg.compute_dtype]
    # bf16: the mid convolutions run on the library -- let it time its algorithms for the
    # (few, fixed) shapes during the eager warm-up steps that precede graph capture.
    # fp32 = parity mode: strict IEEE GEMMs / convolutions and the default heuristics.
    self._tune_convs = self.cd != f32 and bool(cfg.get('cudnn_benchmark', True))
    self._backend_flags()
    self.store = paramlib.ParamStore(cfg, self.device, self.cd, cfg.seed, values)
    self.model = modellib.Model(cfg, self.store)
    self.opt = optim.Optimizer(cfg, self.store)
    self.world = 1
    self.rank = 0
    if torch.distributed.is_available() and torch.distributed.is_initialized():
      self.world = torch.distributed.get_world_size()
      self.rank = torch.distributed.get_rank()
    # bucketed gradient exchange + optimiser under the backward pass (exchange.py): on by default
    # in the bf16 (benchmark) mode and whenever there is more than one rank
    want = cfg.get('grad_buckets', 'auto')
    self.exchange = None
    if self.opt.fused and (want is True or (want == 'auto' and (self.world > 1 or self.cd != f32))):
      from . import exchange as exchangelib
      comm = exchangelib.NcclComm(self.device) if self.world > 1 else None
      self.exchange = exchangelib.GradExchange(self.store, self.opt, comm)
    self.gen = torch.Generator(device=self.device)
    self.gen.manual_seed(cfg.seed * 1000003 + self.rank)     # transform.py:84-85 fold_in(rank)
    self.updates = 0
    self._graphs, self._graph_seen, self._graph_ok = {}, {}, self.opt.fused

  # --------------------------------------------------------------- plugin props
  @property
  def policy_keys(self):
    return '^(enc|dyn|dec|pol)/'

  @property
  def ext_space(self):                                       # dreamerv3/agent.py:88-99
    S = elements.Space
    cfg = self.cfg
    return {'dyn/deter': S(np.float32, (cfg.deter,)),
            'dyn/stoch': S(np.float32, (cfg.stoch, cfg.classes))}

  def init_policy(self, batch_size):                         # agent.py:101-107
    cfg, dev = self.cfg, self.device
    prevact = {
        name: torch.zeros((batch_size, *shape), device=dev,
                          dtype=torch.int32 if kind == 'disc' else f32)
        for name, kind, shape, _ in cfg.actspec}
    return (torch.zeros((batch_size, cfg.deter), dtype=self.cd, device=dev),
            torch.zeros((batch_size, cfg.stoch, cfg.classes), dtype=self.cd, device=dev),
            prevact)

  init_train = init_policy
  init_report = init_policy

  def _backend_flags(self):
    """The library switches are process-wide: every entry point sets what THIS agent's
    mode needs, so agents of both modes can live in one process (the test suite)."""
    if self.cd == f32:
      torch.backends.cuda.matmul.allow_tf32 = False
      torch.backends.cudnn.allow_tf32 = False
    torch.backends.cudnn.benchmark = self._tune_convs

  # --------------------------------------------------------------------- policy
  @torch.no_grad()
  def policy(self, carry, obs, mode='train', noise=None):    # agent.py:115-135
    self._backend_flags()
    cfg, m = self.cfg, self.model
    assert not any(k.startswith('log/') for k in obs), list(obs)          # jax/agent.py:223
    deter, stoch, prevact = carry
    if not isinstance(prevact, dict):                        # a bare tensor: the single action key
      prevact = {self.actkeys[0]: prevact}
    dev = lambda x: x if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x), device=self.device)
    inputs = {k: dev(obs[k]) for k in self.obskeys}
    normalized = None
    if len(cfg.imgkeys) == 1:
      normalized = (getattr(obs, 'normalized', None) or {}).get(cfg.imgkeys[0][0])
    reset = dev(obs['is_first'])
    n = len(reset)
    if noise is None:
      noise = dict(stoch=modellib.gumbel_like((n, cfg.stoch, cfg.classes), self.device, self.gen),
                   action=m.action_noise((n,), self.gen))
    anoise = noise['action']
    if not isinstance(anoise, dict):
      anoise = {self.actkeys[0]: anoise}
    tokens = m.encoder(inputs, normalized)
    (deter, stoch), feat = m.observe(
        (deter, stoch), tokens[:, None], {k: v[:, None] for k, v in prevact.items()}, reset[:, None],
        noise['stoch'][:, None])
    act = m.policy_sample(m.policy_outputs(m.feat2tensor(deter, stoch)), anoise)
    out = {'dyn/deter': deter.to(f32), 'dyn/stoch': stoch.to(f32)}   # _take_outs upcast, jax/agent.py:399-403
    return (deter, stoch, act), dict(act), out

  # ---------------------------------------------------------------------- train
  def _apply_replay_context(self, carry, data):              # agent.py:312-340
    K = self.cfg.replay_context
    deter, stoch, prevact = carry
    if not isinstance(prevact, dict):
      prevact = {self.actkeys[0]: prevact}
    obs = {k: data[k] for k in (*self.obskeys, 'reward', 'is_first', 'is_last', 'is_terminal')}
    act = {k: data[k] for k in self.actkeys}
    if not K:
      pa = {k: torch.cat([prevact[k][:, None], act[k][:, :-1]], 1) for k in act}
      return (deter, stoch), obs, pa, data['stepid']
    first = data['consec'][:, 0] == 0
    rep_deter = data['dyn/deter'][:, K - 1].to(self.cd)
    rep_stoch = data['dyn/stoch'][:, K - 1].to(self.cd)
    # prepend(prev, act)[:, K:] == act[:, K-1:-1] for K >= 1: both branches of agent.py:336-339 agree
    pa = {k: act[k][:, K - 1: -1] for k in act}
    deter = torch.where(first[:, None], rep_deter, deter)
    stoch = torch.where(first[:, None, None], rep_stoch, stoch)
    obs = {k: v[:, K:] for k, v in obs.items()}
    return (deter, stoch), obs, pa, data['stepid'][:, K:]

  def make_noise(self, B, T, out=None):
    """Sampling noise of one train step (an explicit input): Gumbel tensors for the
    categorical latents / actions, normal noise for continuous actions.  With `out` the static
    buffers of the captured step are refilled in place."""
    cfg, dev, H = self.cfg, self.device, self.cfg.imag_length
    shapes = dict(observe=(B, T, cfg.stoch, cfg.classes),
                  imag_stoch=(B * T, H, cfg.stoch, cfg.classes))
    if out is None:
      noise = {k: modellib.gumbel_like(s, dev, self.gen) for k, s in shapes.items()}
      act = self.model.action_noise((B * T, H + 1), self.gen)
      # one scalar discrete action: the bare tensor (the original layout of this dict)
      noise['imag_act'] = act[self.actkeys[0]] if self.model.single_disc else act
      return noise
    for k in shapes:
      modellib.gumbel_(out[k], self.gen)
    for (name, kind, _, _) in cfg.actspec:
      buf = out['imag_act'][name] if isinstance(out['imag_act'], dict) else out['imag_act']
      modellib.gumbel_(buf, self.gen) if kind == 'disc' else buf.normal_(generator=self.gen)
    return out

  # The device work of one update, split where the data-parallel gradient
  # all-reduce sits (embodied/jax/opt.py:52-54): `_fwd_bwd` = replay context,
  # loss, backward into the flat gradient buffer; `_apply` = optimiser chain,
  # slow critic, outputs.  Both are pure stream work (no host reads), so each is
  # captured ONCE into a CUDA graph and replayed: the ~2500 launches of a step
  # cost one cudaGraphLaunch instead of ~64 ms of host enqueue time.
  def _fwd_bwd(self, carry, data, noise):
    carry, obs, prevact, stepid = self._apply_replay_context(carry, data)
    self.store.begin_step()
    self.store.grad.zero_()
    total, carry, outs, metrics = self.model.loss(carry, obs, prevact, noise, update=True)
    if self.exchange is not None:
      self.exchange.begin()
      total.backward()
      metrics = dict(metrics)
      metrics['opt/grad_norm'] = self.exchange.finish()     # all-reduce + optimiser ran underneath
    else:
      total.backward()
    return total, carry, outs, metrics, stepid

  def _apply(self, total, carry, outs, metrics, stepid, data):
    metrics = dict(metrics)
    if self.exchange is None:
      metrics['opt/grad_norm'] = self.opt.launch()
    self.opt.update_slow()
    self.store.begin_step()
    metrics['loss'] = total
    metrics = {k: v.detach() if isinstance(v, torch.Tensor) else v for k, v in metrics.items()}
    feat = outs['feat']
    replay = {'stepid': stepid, 'dyn/deter': feat['deter'].detach().to(f32),
              'dyn/stoch': feat['stoch'].detach().to(f32)}
    carry = (carry[0].detach(), carry[1].detach(), {k: data[k][:, -1].clone() for k in self.actkeys})
    return carry, replay, metrics

  def _allreduce(self):
    if self.world > 1 and self.exchange is None:
      torch.distributed.all_reduce(self.store.grad, op=torch.distributed.ReduceOp.AVG)

  def train(self, carry, data, noise=None):                  # agent.py:137-154, opt.py:31-81
    self._backend_flags()
    mode = self.cfg.get('graph', 'auto')
    if mode and mode != 'off' and self._graph_ok:
      key = tuple((k, tuple(v.shape), v.dtype) for k, v in sorted(data.items()))
      st = self._graphs.get(key)
      if st is None:
        seen = self._graph_seen.get(key, 0)
        self._graph_seen[key] = seen + 1
        if seen >= self.GRAPH_WARMUP:
          st = self._capture(key, carry, data)
      if st is not None:
        return self._train_graphed(st, carry, data, noise)
    return self._train_eager(carry, data, noise)

  GRAPH_WARMUP = 2     # eager steps before capture (cuDNN / cuBLAS / NCCL initialisation)

  def _train_eager(self, carry, data, noise=None):
    B, T = data['is_first'].shape[0], data['is_first'].shape[1] - self.cfg.replay_context
    if noise is None:
      noise = self.make_noise(B, T)
    total, carry, outs, metrics, stepid = self._fwd_bwd(carry, data, noise)
    self._allreduce()
    extra = self.opt.begin_update()
    carry, replay, metrics = self._apply(total, carry, outs, metrics, stepid, data)
    metrics.update(extra)
    self.opt.end_update()
    self.updates += 1
    # detached: a live autograd graph would keep this step's AccumulateGrad
    # nodes (and their stream) alive into a later stream capture
    self.last_outs = _detach(outs)
    return carry, {'replay': replay}, metrics

  def _capture(self, key, carry, data):
    """Capture the two halves of the update for this batch signature.  Returns
    None (and disables graphs, loudly) if the stream capture is refused."""
    import types
    B, T = data['is_first'].shape[0], data['is_first'].shape[1] - self.cfg.replay_context
    st = types.SimpleNamespace()
    st.data = {k: v.clone() for k, v in data.items()}
    st.carry = _clone(tuple(carry))
    st.noise = self.make_noise(B, T)
    scan = self.model.scan
    if scan is not None:
      scan.invalidate()          # the weight packing must be recorded inside the graph
    self.last_outs = None
    pool = torch.cuda.graph_pool_handle()
    st.ga, st.gb = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
    from .. import _lib
    launched = _lib.launch_count()
    try:
      torch.cuda.synchronize()
      with torch.cuda.graph(st.ga, pool=pool):
        mid = self._fwd_bwd(st.carry, st.data, st.noise)
      with torch.cuda.graph(st.gb, pool=pool):
        carry_out, replay, metrics = self._apply(*mid, st.data)
        names = [k for k, v in metrics.items() if isinstance(v, torch.Tensor)]
        mvec = torch.stack([metrics[k].to(f32).reshape(()) for k in names])
      torch.cuda.synchronize()
    except Exception as e:      # noqa: BLE001 -- capture refused: stay on the eager path
      import warnings
      warnings.warn(f'dreamerv3.Agent: CUDA graph capture of the train step failed ({e!r}); '
                    'continuing with eager launches')
      self._graph_ok = False
      self.store.begin_step()
      if scan is not None:
        scan.invalidate()
        from . import scan as scanlib
        if scanlib.GRAPH_TIMERS:         # stopwatches of the refused capture never record
          scanlib.GRAPH_TIMERS.clear()
      try:                               # the failed capture may leave a sticky error behind
        torch.cuda.synchronize()
      except Exception:                  # noqa: BLE001
        pass
      return None
    st.outs, st.carry_out, st.replay = _detach(mid[2]), carry_out, replay
    del mid
    st.names, st.mvec = names, mvec
    st.launches = _lib.launch_count() - launched     # our kernels inside the two graphs
    _lib.launch_count_add((1 << 64) - st.launches)    # recorded, not executed, by the capture
    self._lib = _lib
    if scan is not None:
      scan.invalidate()
    self._graphs[key] = st
    # The capture itself did not execute anything: the caller replays it now.
    return st

  def _train_graphed(self, st, carry, data, noise=None):
    _copy_into(st.carry, tuple(carry))     # carry first: it may alias last step's outputs
    for k, v in st.data.items():
      v.copy_(data[k])
    if noise is None:
      self.make_noise(*st.noise['observe'].shape[:2], out=st.noise)
    else:
      _copy_into(st.noise, noise)
    st.ga.replay()
    self._allreduce()
    extra = self.opt.begin_update()
    st.gb.replay()
    self._lib.launch_count_add(st.launches)
    self.opt.end_update()
    self.updates += 1
    self.last_outs = st.outs
    mv = st.mvec.clone()
    metrics = {k: mv[i] for i, k in enumerate(st.names)}
    metrics.update(extra)
    # stepid: the view of the batch Replay.sample returned (same bytes as the static copy),
    # so that Replay.update recognises its own tensor and needs no device read
    replay = dict(st.replay)
    K = self.cfg.replay_context
    replay['stepid'] = data['stepid'][:, K:] if K else data['stepid']
    return st.carry_out, {'replay': replay}, metrics

  @torch.no_grad()
  def report(self, carry, data, noise=None):                 # agent.py:247-310 (metrics part)
    self._backend_flags()
    carry, obs, prevact, _ = self._apply_replay_context(carry, data)
    B, T = obs['is_first'].shape
    if noise is None:
      noise = self.make_noise(B, T)
    self.store.begin_step()
    total, carry, outs, metrics = self.model.loss(carry, obs, prevact, noise, update=False)
    self.store.begin_step()
    metrics['loss'] = total
    carry = (carry[0], carry[1], {k: data[k][:, -1] for k in self.actkeys})
    return carry, metrics

  def stream(self, st):
    """Batches from Replay.sample are already dense device tensors; the
    reference's device_put + NaN scan (jax/agent.py:326-337) has nothing to do."""
    return st

  def save(self):
    data = self.store.state_dict()
    data['retnorm/lo'] = self.model.ret_lo.cpu().numpy()
    data['retnorm/hi'] = self.model.ret_hi.cpu().numpy()
    data['updates'] = np.asarray(self.updates)
    return data

  def load(self, data, regex=None):
    """`regex`: restore only the parameters whose name matches (embodied/jax/agent.py:343-358,
    run/train.py:88 from_checkpoint_regex); optimiser state and counters of a partial load
    stay as initialised."""
    if regex:
      import re
      pattern = re.compile(regex)
      keep = {k: v for k, v in data.items() if k in self.store.specs and pattern.match(k)}
      if not keep:
        raise KeyError(f'no parameter matches {regex!r}')
      self.store.load_params(keep)
      return
    self.store.load_state_dict(data)
    if 'retnorm/lo' in data:
      self.model.ret_lo.copy_(torch.as_tensor(data['retnorm/lo']))
      self.model.ret_hi.copy_(torch.as_tensor(data['retnorm/hi']))
    self.updates = int(data.get('updates', 0))
    self.opt.sync_count()
