"""The CPU arm of bench.py for the Driver + Replay workloads (BASELINE configs 3 / 4) runs without
a GPU and prints the contract line (`--impl reference`: rank 0 alone works, the line carries
impl / cpu_baseline / e2e with zero copies)."""
import json
import os
import pathlib
import subprocess
import sys

import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent


@pytest.mark.parametrize('workload', ['proprio_rows', 'atari_rows'])
def test_rows_reference_arm_prints_the_contract_line(workload):
  env = dict(os.environ, RANK='0', WORLD_SIZE='1')
  proc = subprocess.run([sys.executable, 'bench.py', '--impl', 'reference', '--workload', workload,
                         '--steps', '1', '--warmup', '1'], cwd=ROOT, env=env, capture_output=True,
                        text=True, timeout=300)
  assert proc.returncode == 0, proc.stderr[-2000:]
  line = json.loads([l for l in proc.stdout.splitlines() if l.startswith('{')][-1])
  assert line['impl'] == 'reference' and line['metric'] == 'env_steps_per_sec' and line['value'] > 0
  assert line['cpu_baseline']['kind'] == 'port' and line['cpu_baseline']['cores'] == 1
  assert line['e2e'] == {'value': line['value'], 'unit': 'env steps/s', 'h2d_bytes_per_step': 0,
                         'd2h_bytes_per_step': 0}


def test_other_ranks_of_the_reference_arm_exit_without_work():
  env = dict(os.environ, RANK='1', WORLD_SIZE='2')
  proc = subprocess.run([sys.executable, 'bench.py', '--impl', 'reference', '--gpus', '2'], cwd=ROOT, env=env,
                        capture_output=True, text=True, timeout=120)
  assert proc.returncode == 0 and proc.stdout.strip() == ''
