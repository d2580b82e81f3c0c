"""bench.py contract checks that run without a GPU: the reference arm (CPU oracle port) prints one JSON line with the
keys the driver reads."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.timeout(600)
def test_reference_arm_json_line():
    env = dict(os.environ, OMP_NUM_THREADS='1')     # torchrun exports this; the arm must still use every core
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '1', '--steps', '1',
                        '--warmup', '0'], capture_output=True, text=True, env=env, cwd=ROOT, timeout=550)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line['impl'] == 'reference' and line['higher_is_better'] is True and line['n_gpus'] == 1
    for key in ('metric', 'value', 'unit', 'steps', 'warmup', 'ms_per_step', 'scaling', 'dtype', 'data', 'config', 'cpu_baseline', 'e2e'):
        assert key in line, key
    assert line['unit'] == 'image-pairs/s' and line['value'] > 0
    assert line['e2e']['h2d_bytes_per_step'] == 0 and line['e2e']['d2h_bytes_per_step'] == 0
    assert line['cpu_baseline']['kind'] == 'port' and line['cpu_baseline']['cores'] >= 1
    if (os.cpu_count() or 1) > 1:
        assert line['cpu_baseline']['cores'] > 1, 'reference arm must not stay on the single thread torchrun grants'


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK='1', WORLD_SIZE='2', LOCAL_RANK='1')
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2', '--steps', '1',
                        '--warmup', '0'], capture_output=True, text=True, env=env, cwd=ROOT, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ''
