"""The bench line committed under profiles/ (the one the DESIGN tables are generated from) carries every
key of the measurement contract: the base keys, `roofline`, `cpu_baseline`, `e2e`, `gpu_launches`,
`clocks`, and this repo's additions (hand-over counters, separate sweeps, per-config stages)."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name):
    with open(os.path.join(ROOT, 'profiles', name)) as f:
        return json.loads(f.read().strip().splitlines()[-1])


def test_single_gpu_line_has_the_contract_keys():
    d = _line('r02_bench_line.json')
    for k in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better',
              'scaling', 'vs_baseline', 'dtype', 'data', 'config', 'roofline', 'cpu_baseline', 'e2e',
              'gpu_launches', 'clocks'):
        assert k in d, k
    assert d['n_gpus'] == 1 and d['higher_is_better'] is True and d['vs_baseline'] is None
    assert d['dtype'] == 'f32' and d['data'] == 'synthetic' and d['unit'] == 'images/s'
    assert 'workload' in d['config'] and 'model' not in d['config']
    r = d['roofline']
    assert r['bound'] == 'hbm' and r['unit'] == 'GB/s'
    assert abs(r['frac'] - r['achieved'] / r['peak']) < 1e-9
    assert r['traffic'] is None or r['traffic'] >= r['algorithmic_bytes_per_launch']
    # value = images of one step / time of one step
    assert abs(d['value'] - d['config']['global_batch'] / (d['ms_per_step'] / 1e3)) < 1e-6 * d['value']
    c = d['cpu_baseline']
    assert c['kind'] in ('reference', 'port') and c['cores'] >= 1 and c['value'] > 0 and c['sample']
    e = d['e2e']
    assert e['unit'] == 'images/s' and e['h2d_bytes_per_step'] > 0 and e['d2h_bytes_per_step'] > 0
    assert 0 < e['value'] < d['value']            # host buffers in, PCIe inside the timed region
    assert d['gpu_launches'] > 0
    assert not set(d['clocks']['reasons']) & {'hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown'}


def test_single_gpu_line_reports_the_hand_over_and_every_baseline_config():
    d = _line('r02_bench_line.json')
    h = d['handoff']
    assert h['produced'] == h['consumed'] > 0 and h['stale'] == 0   # every step's keys came from that step
    assert d['separate_sweeps']['ms_per_step'] > d['ms_per_step']
    names = [c['name'] for c in d['configs']]
    assert [n.split(' ')[0] for n in names] == [f'configs[{i}]' for i in range(5)]
    assert 'eval_step' in d['configs'][3] and 'decode_nms' in d['configs'][0]


def test_multi_gpu_lines_check_parity_in_the_run():
    base = None
    for n in (2, 4, 8):
        d = _line(f'r02_scale_n{n}.json')
        assert d['n_gpus'] == n and d['scaling'] == 'strong' and d['config']['global_batch'] == 256
        p = d['parity_check']
        assert p['ok'] and p['labels_equal'] and p['exchange_status'] == 0
        assert p['sharded_vs_unsharded_rel'] <= p['tolerance_rel'] == 1e-5
        assert d['exchange_status'] == 0
        if base is not None:
            assert d['value'] > base              # more GPUs, more images per second
        base = d['value']
