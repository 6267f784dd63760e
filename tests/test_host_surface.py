"""CPU: the off-path helpers of the reference's function-level surface that this package restates in
torch -- stepfun.sample (camera_utils.py:266) and stepfun.resample (vis.py:144-145) -- against values produced by the
unmodified reference (tests/golden/stepfun_utils.npz, oracle/make_golden.py stepfun_utils)."""
import numpy as np
import pytest
import torch

from refnerf_pl_b200 import stepfun
from tests._cases import GOLDEN


@pytest.fixture(scope='module')
def gold():
    return np.load(GOLDEN + '/stepfun_utils.npz')


def test_sample_matches_reference(gold):
    out = stepfun.sample(torch.tensor(gold['cam_theta']), torch.log(torch.tensor(gold['cam_lengths'])), 31)
    assert out.dtype == torch.float64 and np.array_equal(out.numpy(), gold['cam_sample'])
    t, wl = torch.tensor(gold['s_t']), torch.tensor(gold['s_logits'])
    assert np.array_equal(stepfun.sample(t, wl, 16).numpy(), gold['s_lin'])
    assert np.array_equal(stepfun.sample(t, wl, 16, deterministic_center=True).numpy(), gold['s_center'])
    # sorted input -> sorted samples inside the support
    s = stepfun.sample(t, wl, 64)
    assert bool((s[:, 1:] >= s[:, :-1]).all()) and float(s.min()) >= float(t.min()) and float(s.max()) <= float(t.max())


@pytest.mark.parametrize('use_avg', [False, True])
def test_resample_matches_reference(gold, use_avg):
    t, tp = torch.tensor(gold['r_t']), torch.tensor(gold['r_tp'])
    for key, v in (('3', torch.tensor(gold['r_v3'])), ('1', torch.tensor(gold['r_v1']))):
        out = stepfun.resample(t, tp, v, use_avg=use_avg)
        ref = gold[f'r_out{key}_{int(use_avg)}']
        assert out.shape == ref.shape and out.dtype == torch.float64
        assert np.abs(out.numpy() - ref).max() <= 1e-13
    # re-binning by sum preserves the integral when t covers the support of tp
    v = torch.tensor(gold['r_v1'])
    assert abs(float(stepfun.resample(t, tp, v).sum()) - float(v.sum())) <= 1e-9
