"""On-disk replay format: `{time}-{uuid}-{succ}-{length}.npz` chunk files
(embodied/core/chunk.py:29-33,64-99) and Replay.save / load (replay.py:295-359).

* tests/golden/ref_chunks/ was WRITTEN BY THE REFERENCE's own Replay.save
  (oracle/gen_golden.py gen_chunkdir); the product loads it and must then sample exactly what
  the reference sampled after loading the same directory (ref_chunks_expected.npz).
* Where /root/reference exists, the other direction runs live: the product writes a directory,
  the reference's own Replay.load reads it, and both buffers sample identical bytes."""
import pathlib
import shutil

import numpy as np
import pytest

import embodied_b200 as embodied
from embodied_b200 import elements
from oracle import gen_golden, refload
import doubles
import golden_cases

needs_ref = pytest.mark.skipif(
    not refload.available(), reason='/root/reference not on this machine')
SPEC = gen_golden.CHUNKDIR_SPEC


def product(directory, **kw):
  return embodied.Replay(
      SPEC['length'], SPEC['capacity'], chunksize=SPEC['chunksize'], directory=str(directory),
      seed=SPEC['seed'], store=doubles.HostStore(SPEC['chunksize'], staging_rows=4), **kw)


def fields(name):
  time, uuid, succ, length = pathlib.Path(name).stem.split('-')
  return uuid, succ, int(length)


def test_product_loads_directory_written_by_the_reference(tmp_path):
  shutil.copytree(golden_cases.GOLDEN / 'ref_chunks', tmp_path / 'chunks')
  want = np.load(golden_cases.GOLDEN / 'ref_chunks_expected.npz')
  replay = product(tmp_path / 'chunks')
  replay.load()
  assert len(replay) == int(want['len'])
  for i in range(3):
    golden_cases.check_batch(replay.sample(SPEC['batch']), None, f'sample{i}/', want)


def test_product_writes_the_same_chunks_as_the_reference(tmp_path):
  """Same seeded stream, same save() calls: same set of (uuid, succ, length) file names and the
  same arrays inside (time stamps aside)."""
  replay = product(tmp_path, save_wait=True)
  gen_golden.fill_chunkdir_stream(replay, SPEC)
  ours = {fields(p.name): p for p in tmp_path.glob('*.npz')}
  theirs = {fields(p.name): p for p in (golden_cases.GOLDEN / 'ref_chunks').glob('*.npz')}
  assert sorted(ours) == sorted(theirs)
  for key in theirs:
    with np.load(ours[key]) as a, np.load(theirs[key]) as b:
      assert sorted(a.keys()) == sorted(b.keys())
      for k in b.keys():
        assert a[k].dtype == b[k].dtype and a[k].tobytes() == b[k].tobytes(), (key, k)


@needs_ref
def test_reference_loads_directory_written_by_the_product(tmp_path):
  ns = refload.load()
  ns.elements.UUID.reset(debug=True)
  try:
    writer = product(tmp_path, save_wait=True)
    gen_golden.fill_chunkdir_stream(writer, SPEC)
    theirs = ns.replay.Replay(length=SPEC['length'], capacity=SPEC['capacity'],
                              chunksize=SPEC['chunksize'], directory=str(tmp_path), seed=SPEC['seed'])
    theirs.load()
    ours = product(tmp_path)
    ours.load()
    assert len(theirs) == len(ours) > 0
    for _ in range(4):
      a, b = ours.sample(SPEC['batch']), theirs.sample(SPEC['batch'])
      assert sorted(a) == sorted(b)
      for k in b:
        assert golden_cases.tonp(a[k]).tobytes() == np.asarray(b[k]).tobytes(), k
  finally:
    ns.elements.UUID.reset(debug=False)


def test_load_rejects_chunks_longer_than_the_slab(tmp_path):
  shutil.copytree(golden_cases.GOLDEN / 'ref_chunks', tmp_path / 'chunks')
  small = embodied.Replay(
      SPEC['length'], SPEC['capacity'], chunksize=4, directory=str(tmp_path / 'chunks'),
      store=doubles.HostStore(4, staging_rows=4))
  with pytest.raises(ValueError, match='chunksize'):
    small.load()
