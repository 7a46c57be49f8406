"""The slot tables behind emb_pack_tiles (embodied_b200/dreamerv3/scan.py): a numpy
walk of the tables -- exactly what csrc/pack.cu does per lane -- must reproduce the
op-by-op torch formulation (pack_matrix) of every in-scan weight matrix
(dreamerv3/rssm.py:135-159, 81-86), forward and transposed, for both bf16 layouts."""
import numpy as np
import pytest
import torch

from embodied_b200.dreamerv3 import config as C, params as P, scan as S


def walk_tables(store, name, ksteps, segments, engine, ncta, unit=1, groups=1):
  per, tabs = S.slot_tables(segments, engine, ncta, unit, groups)
  master = store.master.detach().numpy()
  dst = np.zeros((ncta, ksteps, per, 8, 4, 2, 2), np.float32)
  for first, count, off, ks, ns in tabs:
    k = np.arange(count * 16, dtype=np.int64)
    idx = (off[:, None, None] + k[None, :, None] * ks[:, None, None].astype(np.int64) +
           np.arange(8)[None, None, :] * ns[:, None, None].astype(np.int64))
    vals = np.where(off[:, None, None] >= 0, master[np.maximum(idx, 0)], 0.0)   # (slots, K, nn)
    vals = vals.reshape(ncta, per, count, 2, 4, 2, 8)      # cta, tile, kstep, reg, kq, half, nn
    dst[:, first:first + count] = vals.transpose(0, 2, 1, 6, 4, 3, 5)
  return torch.from_numpy(dst).to(torch.bfloat16)


@pytest.mark.parametrize('size', ['size1m', 'size12m'])
@pytest.mark.parametrize('engine', [S.ENG_BF16, S.ENG_LEGACY])
@pytest.mark.parametrize('ncta', [148, 20])
def test_slot_tables_reproduce_the_torch_packing(size, engine, ncta):
  cfg = C.make(size)
  store = P.ParamStore(cfg, 'cpu', torch.float32, 3)
  S.FUSED_PACK = False
  try:
    want = S.pack(store, cfg, engine, ncta)
    want_bwd = S.pack_bwd(store, cfg, engine, ncta)
  finally:
    S.FUSED_PACK = True
  got = S._pack_fused(store, cfg, engine, ncta, pack_tiles=walk_tables)
  got_bwd = S._pack_bwd_fused(store, cfg, engine, ncta, pack_tiles=walk_tables)
  assert set(got) == {'w_ph1', 'w_logit', 'w_hid', 'w_gru'}
  assert set(got_bwd) == {'wt_in1', 'wt_logit', 'wt_ph1', 'wt_gru', 'wt_hid'}
  for name, x in {**got, **got_bwd}.items():
    ref = {**want, **want_bwd}[name]
    assert x.shape == ref.shape, (name, x.shape, ref.shape)
    assert torch.equal(x, ref), name
