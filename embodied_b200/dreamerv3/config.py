"""dreamerv3 hyper-parameters (dreamerv3/configs.yaml:81-145), flattened.

`make(size='size200m', **overrides)`; the size presets are the reference's
regex blocks (configs.yaml:120-145: `.*\\.rssm`, `.*\\.depth`, `.*\\.units`).
"""

SIZES = {
    'size1m': dict(deter=512, hidden=64, classes=4, depth=4, units=64),
    'size12m': dict(deter=2048, hidden=256, classes=16, depth=16, units=256),
    'size25m': dict(deter=3072, hidden=384, classes=24, depth=24, units=384),
    'size50m': dict(deter=4096, hidden=512, classes=32, depth=32, units=512),
    'size100m': dict(deter=6144, hidden=768, classes=48, depth=48, units=768),
    'size200m': dict(deter=8192, hidden=1024, classes=64, depth=64, units=1024),
    'size400m': dict(deter=12288, hidden=1536, classes=96, depth=96, units=1536),
}


class Config(dict):
  __getattr__ = dict.__getitem__

  def update(self, *a, **kw):
    super().update(*a, **kw)
    return self


def make(size='size200m', **over):
  cfg = Config(
      # agent.dyn.rssm
      deter=8192, hidden=1024, stoch=32, classes=64, blocks=8, unimix=0.01,
      free_nats=1.0, imglayers=2, obslayers=1, dynlayers=1,
      # agent.enc.simple / agent.dec.simple
      depth=64, mults=(2, 3, 4, 4), kernel=5, units=1024, bspace=8,
      # heads
      bins=255, rew_layers=1, con_layers=1, pol_layers=3, val_layers=3,
      # imagination / losses
      imag_length=15, horizon=333, contdisc=True, lam=0.95, actent=3e-4,
      slowreg=1.0, slowrate=0.02, replay_context=1,
      retnorm_rate=0.01, retnorm_limit=1.0, perclo=5.0, perchi=95.0,
      scales=dict(image=1.0, rew=1.0, con=1.0, dyn=1.0, rep=0.1,
                  policy=1.0, value=1.0, repval=0.3),
      # agent.opt
      lr=4e-5, agc=0.3, eps=1e-20, beta1=0.9, beta2=0.999, warmup=1000,
      pmin=1e-3,
      # spaces (filled by the Agent from obs_space / act_space)
      image=(64, 64, 3), actions=5,
      # runtime (the `jax:` block's role): compute dtype and seed
      compute_dtype='bfloat16', seed=0,
  )
  cfg.update(SIZES[size])
  cfg.update(over)
  return cfg
