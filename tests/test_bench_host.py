"""CPU: the host-side pieces of bench.py — algorithmic byte counts, captured batches with the reference
loader's semantics, the bounded CPU baselines, the clock sampler's degradation without nvidia-smi."""
import os
import stat
import sys

import numpy as np

from hsk_testutil import ROOT

sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_algorithmic_bytes_match_survey_figures():
    ab = bench.algorithmic_bytes(6040, 3706, 402, 8192, 50)          # cfg2
    assert ab['A'] == 4 * 402 * 8192 * 52 == 684_982_272              # "A 685 MB"
    assert ab['P'] == (6040 + 3706) * 402 + 3706 == 3_921_598
    assert ab['total'] == ab['gather_scatter'] + ab['adamw'] and ab['adamw'] == 28 * ab['P']
    assert abs(ab['total'] - 1_486.5e6) < 0.1e6                              # "1 486.5 MB/step"
    ab3 = bench.algorithmic_bytes(69878, 10677, 128, 8192, 100)       # cfg3
    assert ab3['P'] == 10_321_717 and abs(ab3['total'] - 1_157.9e6) < 0.1e6


def test_make_batches_negatives_are_never_train_items():
    from hassaku_b200.data.synthetic import make_interactions
    data = make_interactions(300, 200, 6000, seed=0)
    us, its = bench.make_batches(data, 256, 7, 3, seed=1)
    tr = data.train.tocsr()
    assert len(us) == 3 and us[0].shape == (256,) and its[0].shape == (256, 8) and its[0].dtype == np.int64
    for u, i in zip(us, its):
        assert all(tr[uu, ii[0]] != 0 for uu, ii in zip(u, i))                  # column 0 is a train positive
        assert not any(tr[uu, j] != 0 for uu, ii in zip(u, i) for j in ii[1:])  # dataloader.py:112-120 rejection
    us2, its2 = bench.make_batches(data, 256, 7, 3, seed=1)
    assert all(np.array_equal(a, b) for a, b in zip(its, its2))                 # seeded


def test_cpu_baselines_report_the_contract_keys():
    from hassaku_b200.data.synthetic import make_interactions
    data = make_interactions(300, 200, 6000, seed=0)
    wl = ('tiny', 16, 64, 5, 'bpr', 3e-4, 4e-5)
    us, its = bench.make_batches(data, 64, 5, 2)
    r = bench.cpu_baseline(wl, data, us, its, budget_s=0.5)
    assert r['unit'] == 'triples/s' and r['kind'] == 'port' and r['value'] > 0 and r['cores'] >= 1 and 'sample' in r
    e = bench.cpu_eval_baseline(wl, data, n_users=64, eval_batch_size=32)
    assert e['unit'] == 'users/s' and e['kind'] == 'port' and e['value'] > 0 and 0 <= e['ndcg@10_of_sample'] <= 1


def test_clock_sampler_degrades_and_parses(tmp_path, monkeypatch):
    monkeypatch.setenv('PATH', str(tmp_path))                      # no nvidia-smi at all
    with bench.ClockSampler(0) as c:
        c.wait_ready(timeout=1.0)
    assert c.summary() == {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
    fake = tmp_path / 'nvidia-smi'
    fake.write_text('#!/bin/bash\nwhile true; do echo "1965, 1965, 400.1, Not Active, Not Active, Not Active, Active"; '
                    'echo "1200, 1965, 90.0, Not Active, Not Active, Not Active, Not Active"; sleep 0.05; done\n')
    fake.chmod(fake.stat().st_mode | stat.S_IEXEC)
    monkeypatch.setenv('PATH', str(tmp_path) + os.pathsep + '/usr/bin' + os.pathsep + '/bin')
    with bench.ClockSampler('0,1') as c:
        c.wait_ready(timeout=5.0)
        assert len(c.rows) >= 1                                    # ready means: the first sample has arrived
    s = c.summary()
    assert s['sm_mhz'] == 1965.0 and s['sm_max_mhz'] == 1965.0 and s['reasons'] == ['sw_power_cap'] and s['samples'] >= 2
    with bench.ClockSampler('0', enabled=False) as c:              # ranks other than 0 in the sharded bench
        c.wait_ready()
    assert c.summary()['samples'] == 0
