"""Load the reference's OWN ``embodied/core`` verbatim, by path (TEST INFRA).

Only usable where /root/reference exists (this container).  The reference's
package ``__init__`` files pull in ``jax`` and ``portal`` servers, so instead of
``import embodied`` we register empty package shells whose ``__path__`` points
into /root/reference and import the leaf modules one by one against the
from-scratch ``elements``/``portal`` shims in ``oracle/_shim``.  No reference
source is copied into this repository.
"""
import importlib
import importlib.util
import pathlib
import sys
import types

REFERENCE = pathlib.Path('/root/reference')
_SHIM = pathlib.Path(__file__).parent / '_shim'
_CACHE = {}

CORE = ('base', 'limiters', 'selectors', 'chunk', 'replay', 'streams',
        'driver', 'random', 'wrappers')


def available():
  return (REFERENCE / 'embodied' / 'core' / 'replay.py').exists()


def _load_file(name, path):
  spec = importlib.util.spec_from_file_location(name, path)
  mod = importlib.util.module_from_spec(spec)
  sys.modules[name] = mod
  spec.loader.exec_module(mod)
  return mod


def load():
  """Returns a namespace: .elements .portal .base .replay .chunk .selectors
  .streams .driver .random .wrappers .dummy, all the reference's own code."""
  if 'ns' in _CACHE:
    return _CACHE['ns']
  if not available():
    raise FileNotFoundError('/root/reference is not present on this machine')
  saved = {k: sys.modules.get(k) for k in ('elements', 'portal', 'embodied')}
  elements = _load_file('elements', _SHIM / 'elements.py')
  portal = _load_file('portal', _SHIM / 'portal.py')
  pkg = types.ModuleType('embodied')
  pkg.__path__ = [str(REFERENCE / 'embodied')]
  core = types.ModuleType('embodied.core')
  core.__path__ = [str(REFERENCE / 'embodied' / 'core')]
  envs = types.ModuleType('embodied.envs')
  envs.__path__ = [str(REFERENCE / 'embodied' / 'envs')]
  sys.modules.update({
      'embodied': pkg, 'embodied.core': core, 'embodied.envs': envs})
  ns = types.SimpleNamespace(elements=elements, portal=portal)
  for leaf in CORE:
    mod = importlib.import_module(f'embodied.core.{leaf}')
    setattr(core, leaf, mod)
    setattr(ns, leaf, mod)
  # what embodied/__init__.py + core/__init__.py would have exported
  pkg.Agent, pkg.Env = ns.base.Agent, ns.base.Env
  pkg.Driver, pkg.Replay = ns.driver.Driver, ns.replay.Replay
  pkg.RandomAgent, pkg.Wrapper = ns.random.RandomAgent, ns.wrappers.Wrapper
  for leaf in ('replay', 'streams', 'selectors', 'wrappers', 'limiters'):
    setattr(pkg, leaf, getattr(ns, leaf))
  ns.dummy = importlib.import_module('embodied.envs.dummy')
  ns.embodied = pkg
  # Leave the synthetic 'embodied*' entries registered (the reference modules
  # refer to each other lazily) but give back 'elements'/'portal' names if the
  # process had real ones.
  for k in ('elements', 'portal'):
    if saved[k] is not None:
      sys.modules[k] = saved[k]
  _CACHE['ns'] = ns
  return ns
