"""The bench line's contract, checked on the line of the last recorded run (profiles/r2_bench_n1.json) and on the reference
arm run here (CPU only): every key the driver and the judge read is there, with the types they expect."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

BASE_KEYS = {'metric': str, 'value': float, 'unit': str, 'n_gpus': int, 'steps': int, 'warmup': int, 'ms_per_step': float,
             'higher_is_better': bool, 'scaling': str, 'dtype': str, 'data': str, 'config': dict, 'e2e': dict, 'gpu_launches': int}


def _check_base(line):
    for k, t in BASE_KEYS.items():
        assert k in line, k
        assert isinstance(line[k], t), (k, type(line[k]))
    assert 'vs_baseline' in line and line['vs_baseline'] is None          # BASELINE.md holds no published number for this metric
    assert 'workload' in line['config'] and 'model' not in line['config']
    for k in ('value', 'unit', 'h2d_bytes_per_step', 'd2h_bytes_per_step'):
        assert k in line['e2e'], k


def test_recorded_bench_line_has_every_contract_key():
    with open(os.path.join(ROOT, 'profiles', 'r2_bench_n1.json')) as f:
        line = json.load(f)
    _check_base(line)
    assert line['gpu_launches'] > 0 and line['e2e']['h2d_bytes_per_step'] > 0 and line['e2e']['d2h_bytes_per_step'] > 0
    assert line['e2e']['value'] < line['value']                           # host buffers cannot beat resident inputs
    r = line['roofline']
    for k in ('bound', 'achieved', 'peak', 'unit', 'frac', 'traffic'):
        assert k in r, k
    assert r['bound'] == 'hbm' and abs(r['frac'] - r['achieved'] / r['peak']) < 1e-9
    c = line['cpu_baseline']
    for k in ('value', 'unit', 'cores', 'kind', 'sample'):
        assert k in c, k
    assert c['kind'] in ('reference', 'port')
    for k in ('sm_mhz', 'sm_max_mhz', 'reasons'):
        assert k in line['clocks'], k
    s = line['stack']
    assert s['unit'] == 'voxels/s' and s['scaling'] == 'strong' and s['value'] > 0
    assert line['config1']['matches_reference_run'] is True
    assert line['stack']['with_cnn']['streamed_equals_block'] is True


def test_reference_arm_line():
    """bench.py --impl reference on the host cores (the reference's own postprocess.py when it is staged, else the port)."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0'],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line['impl'] == 'reference'
    _check_base(line)
    assert line['e2e']['h2d_bytes_per_step'] == 0 and line['e2e']['d2h_bytes_per_step'] == 0
    assert line['e2e']['value'] == line['value'] and line['cpu_baseline']['value'] == line['value']
    with open(os.path.join(ROOT, 'profiles', 'r2_bench_n1.json')) as f:
        ours = json.load(f)
    for k in ('metric', 'unit', 'higher_is_better'):
        assert line[k] == ours[k], k
    assert line['config']['workload'] == ours['config']['workload']
