"""Host-side index logic of embodied_b200.Replay (numpy storage double, no
GPU): golden vectors of the real reference, and the invariants of the
reference's tests/test_replay.py ported through a 5-line dataset() adapter
(Replay.dataset was removed upstream)."""
import collections
import threading
import time

import numpy as np
import pytest

import embodied_b200 as embodied
from embodied_b200 import elements
from oracle import gen_golden
import doubles
import golden_cases


def make(length, capacity=None, chunksize=1024, online=False, seed=0, **kw):
  return embodied.Replay(
      length, capacity, chunksize=chunksize, online=online, seed=seed,
      store=doubles.HostStore(chunksize, staging_rows=kw.pop('staging_rows', 16)),
      **kw)


def dataset(replay, batch=1):
  stream = embodied.streams.Stateless(replay.sample, batch)
  while True:
    yield {k: golden_cases.tonp(v)[0] for k, v in next(stream).items()}


@pytest.mark.parametrize('name', sorted(gen_golden.REPLAY_CASES))
def test_product_index_logic_matches_reference_golden(name):
  golden_cases.run_replay_case(name, make)


@pytest.mark.parametrize('name', sorted(gen_golden.REPLAY_CASES))
def test_add_batch_equals_row_by_row(name):
  def adder(replay, rows):
    replay.add_batch({k: np.stack([r[k] for r in rows]) for k in rows[0]})
  golden_cases.run_replay_case(name, make, adder)


def test_multiple_keys():
  replay = make(length=5, capacity=10)
  for step in range(30):
    replay.add({'image': np.zeros((64, 64, 3)), 'action': np.zeros(12)})
  seq = next(dataset(replay))
  assert set(seq.keys()) == {'stepid', 'image', 'action'}
  assert seq['stepid'].shape == (5, 20)
  assert seq['image'].shape == (5, 64, 64, 3)
  assert seq['action'].shape == (5, 12)


@pytest.mark.parametrize(
    'length,workers,capacity',
    [(2, 1, 2), (5, 1, 10), (1, 2, 2), (5, 3, 15), (2, 7, 20)])
def test_capacity_exact(length, workers, capacity):
  replay = make(length, capacity)
  for step in range(30):
    for worker in range(workers):
      replay.add({'step': step}, worker)
    target = min(workers * max(0, (step + 1) - length + 1), capacity)
    assert len(replay) == target


@pytest.mark.parametrize(
    'length,workers,capacity,chunksize',
    [(2, 1, 2, 128), (5, 1, 10, 128), (1, 2, 2, 128),
     (5, 3, 15, 128), (2, 7, 20, 128), (7, 2, 27, 4)])
def test_sample_sequences(length, workers, capacity, chunksize):
  replay = make(length, capacity, chunksize=chunksize)
  for step in range(30):
    for worker in range(workers):
      replay.add({'step': step, 'worker': worker}, worker)
  ds = dataset(replay)
  for _ in range(10):
    seq = next(ds)
    assert (seq['step'] - seq['step'][0] == np.arange(length)).all()
    assert (seq['worker'] == seq['worker'][0]).all()


@pytest.mark.parametrize(
    'length,capacity', [(1, 1), (2, 2), (5, 10), (1, 2), (5, 15), (2, 20)])
def test_sample_single(length, capacity):
  replay = make(length, capacity)
  for step in range(length):
    replay.add({'step': step})
  ds = dataset(replay)
  for _ in range(10):
    assert (next(ds)['step'] == np.arange(length)).all()


def test_sample_uniform():
  replay = make(capacity=20, length=5, seed=0)
  for step in range(7):
    replay.add({'step': step})
  assert len(replay) == 3
  histogram = collections.defaultdict(int)
  ds = dataset(replay)
  for _ in range(100):
    histogram[int(next(ds)['step'][0])] += 1
  assert len(histogram) == 3, histogram
  assert all(v > 20 for v in histogram.values())


def test_workers_simple():
  replay = make(length=2, capacity=20)
  replay.add({'step': 0}, worker=0)
  replay.add({'step': 1}, worker=1)
  replay.add({'step': 2}, worker=0)
  replay.add({'step': 3}, worker=1)
  ds = dataset(replay)
  for _ in range(10):
    assert tuple(next(ds)['step']) in ((0, 2), (1, 3))


def test_workers_random(length=4, capacity=30):
  rng = np.random.default_rng(seed=0)
  replay = make(length, capacity)
  streams = {i: iter(range(10)) for i in range(3)}
  for _ in range(40):
    worker = int(rng.integers(0, 3, ()))
    try:
      replay.add({'step': next(streams[worker]), 'stream': worker}, worker)
    except StopIteration:
      pass
  histogram = collections.defaultdict(int)
  ds = dataset(replay)
  for _ in range(10):
    seq = next(ds)
    assert (seq['step'] - seq['step'][0] == np.arange(length)).all()
    assert (seq['stream'] == seq['stream'][0]).all()
    histogram[int(seq['stream'][0])] += 1
  assert all(count > 0 for count in histogram.values())


def test_slab_reuse_never_aliases_within_a_launch():
  # tiny capacity: every add evicts; HostStore asserts no duplicate rows per
  # launch, and data must stay intact across slab recycling.
  replay = make(3, 4, chunksize=2, staging_rows=8)
  for step in range(200):
    replay.add({'step': np.int64(step)})
    if step > 10 and step % 3 == 0:
      seq = replay.sample(2)['step']
      assert (np.diff(seq, axis=1) == 1).all()
  assert replay.store.nslabs <= 16, replay.store.nslabs


def test_update_roundtrip_and_evicted_rows_skipped():
  replay = make(4, 6, chunksize=3)
  for step in range(12):
    replay.add({'step': np.int32(step), 'lat': np.zeros(2, np.float32)})
  batch = replay.sample(3)
  old_ids = batch['stepid'].copy()
  new = np.arange(3 * 4 * 2, dtype=np.float32).reshape(3, 4, 2) + 1
  replay.update({'stepid': batch['stepid'], 'lat': new})
  again = {tuple(r[0]): l for r, l in zip(batch['stepid'], new)}
  for _ in range(20):
    got = replay.sample(1)
    key = tuple(got['stepid'][0, 0])
    if key in again:
      assert (got['lat'][0] == again[key]).all()
  for step in range(12, 40):   # evict everything that was sampled
    replay.add({'step': np.int32(step), 'lat': np.zeros(2, np.float32)})
  replay.update({'stepid': old_ids, 'lat': new + 100})   # silently skipped
  for _ in range(20):
    assert (replay.sample(1)['lat'] == 0).all()


# ----------------------------------------------------------- save / load

@pytest.mark.parametrize(
    'length,capacity,chunksize',
    [(3, 10, 128), (5, 100, 128), (5, 25, 2)])
def test_restore_exact(tmpdir, length, capacity, chunksize):
  replay = make(length, capacity, chunksize=chunksize, directory=tmpdir,
                save_wait=True)
  for step in range(30):
    replay.add({'step': step})
  num_items = np.clip(30 - length + 1, 0, capacity)
  assert len(replay) == num_items
  data = replay.save()
  replay = make(length, capacity, directory=tmpdir)
  replay.load(data)
  assert len(replay) == num_items
  ds = dataset(replay)
  for _ in range(len(replay)):
    seq = next(ds)['step']
    assert len(seq) == length and (np.diff(seq) == 1).all()


@pytest.mark.parametrize('workers', [1, 2, 5])
@pytest.mark.parametrize(
    'length,capacity,chunksize', [(3, 10, 5), (5, 100, 12)])
def test_restore_chunks_workers(tmpdir, workers, length, capacity, chunksize):
  capacity *= workers
  replay = make(length, capacity, chunksize=chunksize, directory=tmpdir,
                save_wait=True)
  for step in range(50):
    for worker in range(workers):
      replay.add({'step': step}, worker)
  num_items = np.clip((50 - length + 1) * workers, 0, capacity)
  assert len(replay) == num_items
  data = replay.save()
  filenames = list(elements.Path(tmpdir).glob('*.npz'))
  lengths = [int(x.stem.split('-')[3]) for x in filenames]
  stored_steps = min(capacity // workers + length - 1, 50)
  total_chunks = int(np.ceil(50 / chunksize))
  pruned_chunks = int(np.floor((50 - stored_steps) / chunksize))
  assert len(filenames) == (total_chunks - pruned_chunks) * workers
  last_chunk_empty = total_chunks * chunksize - 50
  saved_steps = (total_chunks - pruned_chunks) * chunksize - last_chunk_empty
  assert sum(lengths) == saved_steps * workers
  replay = make(length, capacity, chunksize=chunksize, directory=tmpdir)
  replay.load(data)
  assert len(replay) == num_items
  ds = dataset(replay)
  for _ in range(len(replay)):
    assert len(next(ds)['step']) == length


@pytest.mark.parametrize(
    'length,capacity,chunksize', [(3, 10, 128), (5, 100, 128), (5, 25, 2)])
def test_restore_insert(tmpdir, length, capacity, chunksize):
  replay = make(length, capacity, chunksize=chunksize, directory=tmpdir,
                save_wait=True)
  inserts = int(1.5 * chunksize)
  for step in range(inserts):
    replay.add({'step': step})
  num_items = np.clip(inserts - length + 1, 0, capacity)
  assert len(replay) == num_items
  data = replay.save()
  replay = make(length, capacity, directory=tmpdir)
  replay.load(data)
  assert len(replay) == num_items
  for step in range(inserts):
    replay.add({'step': step})
  num_items = np.clip(2 * (inserts - length + 1), 0, capacity)
  assert len(replay) == num_items


def test_threading(tmpdir, length=5, capacity=128, chunksize=32,
                   adders=8, samplers=4):
  replay = make(length, capacity, chunksize=chunksize, directory=tmpdir,
                save_wait=True)
  running = [True]
  errors = []

  def adder():
    ident = threading.get_ident()
    step = 0
    while running[0]:
      replay.add({'step': step}, worker=ident)
      step += 1
      time.sleep(0.001)

  def sampler():
    try:
      ds = dataset(replay)
      while running[0]:
        seq = next(ds)
        assert (seq['step'] - seq['step'][0] == np.arange(length)).all()
        time.sleep(0.001)
    except Exception as e:
      errors.append(e)
      raise

  workers = [threading.Thread(target=adder) for _ in range(adders)]
  workers += [threading.Thread(target=sampler) for _ in range(samplers)]
  try:
    [w.start() for w in workers]
    for _ in range(3):
      time.sleep(0.1)
      stats = replay.stats()
      assert stats['inserts'] > 0
      assert stats['samples'] > 0
      data = replay.save()
      time.sleep(0.1)
      replay.load(data)
  finally:
    running[0] = False
    [w.join() for w in workers]
  assert not errors, errors
  assert len(replay) == capacity


def test_update_overlapping_windows_last_writer_wins():
  replay = make(3, 50, chunksize=4)
  for step in range(6):
    replay.add({'step': np.int32(step), 'lat': np.zeros(1, np.float32)})
  batch = replay.sample(8)          # 4 items, 8 draws: overlaps guaranteed
  new = np.arange(8 * 3, dtype=np.float32).reshape(8, 3, 1) + 1
  replay.update({'stepid': batch['stepid'], 'lat': new})
  want = {}
  for b in range(8):
    for t in range(3):
      want[int(batch['step'][b, t])] = new[b, t, 0]
  got = replay.sample(8)
  for b in range(8):
    for t in range(3):
      assert got['lat'][b, t, 0] == want[int(got['step'][b, t])]
  kinds = [k for k, _ in replay.store.launches]
  assert kinds.count('scatter') == 1


def test_shift_crosses_chunk_completed_early_by_save(tmp_path):
  """save() completes the current chunk early (length < size, successor set); a window
  offset that crosses it must land in the successor, not past the chunk's length
  (Replay._shift used chunk.size)."""
  replay = make(6, 100, chunksize=32, directory=str(tmp_path), save_wait=True)
  for step in range(20):
    replay.add({'step': np.int32(step), 'lat': np.zeros(1, np.float32)})
  replay.save()
  for step in range(20, 40):
    replay.add({'step': np.int32(step), 'lat': np.zeros(1, np.float32)})
  (c0, _) = replay.items[0]
  chunk = replay.chunks[c0]
  assert chunk.length == 20 and chunk.succ != elements.UUID(0)
  assert replay._shift((c0, 15), 4) == (c0, 19)
  assert replay._shift((c0, 15), 5) == (chunk.succ, 0)
  assert replay._shift((c0, 15), 6) == (chunk.succ, 1)
  rows = replay._rows_of(*replay._shift((c0, 15), 6), 2)
  assert len(rows) == 2
  # a chunk still being written (no successor) keeps the slab size as its room
  cur = replay.chunks[chunk.succ]
  assert cur.succ == elements.UUID(0)
  assert replay._shift((cur.uuid, 3), 10) == (cur.uuid, 13)
