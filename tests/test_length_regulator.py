"""length regulator + duration rounding (integer frame indexing)."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from tts_sambert_hifigan_b200 import synth
import oracle.length_regulator as olr

B, Tph, D = 4, 37, 24


def _inputs():
    henc = synth.normal(201, (B, Tph, D))
    g = load_golden("length_regulator")
    return henc, g["dur"], g


def test_oracle_matches_reference_golden():
    henc, dur, g = _inputs()
    assert np.array_equal(olr.length_regulate(henc, dur), g["hlr"])
    log_dur = synth.uniform(203, (B, 400), 3.0)
    assert np.array_equal(olr.durations_from_log(log_dur), g["dur_from_log"])     # 0 of 1600 differ on this fixture
    # a fixture built to sit ON ties: log(k + 0.5) rounded to float32, for k = 1 .. 4000.  The live reference
    # (torch CPU vector expf, 1 ulp) and the correctly rounded exponential may part only where exp(x) is within
    # one float32 ulp of k + 0.5 -- and nowhere else
    ties = np.log(np.arange(1, 4001, dtype=np.float64) + 0.5).astype(np.float32)
    import torch
    ref = torch.clamp(torch.exp(torch.from_numpy(ties)).round().long(), min=1).numpy()  # reference :746-748, same ATen ops
    mine = olr.durations_from_log(ties)
    diff = np.flatnonzero(ref != mine)
    x = np.exp(ties.astype(np.float64))
    ulp = np.spacing(x.astype(np.float32)).astype(np.float64)
    assert all(abs(x[i] - (np.floor(x[i]) + 0.5)) <= ulp[i] for i in diff)
    print(f"tie fixture: {len(diff)} of {len(ties)} durations differ between ATen's CPU expf and the "
          f"correctly rounded float32 exponential (all within 1 ulp of a tie)")


def test_oracle_reference_known_answer():
    # reference tests/test_length_regulator.py:70-104 (repeat logic known-answer test)
    henc = np.array([[[1, 2, 3, 4], [5, 6, 7, 8], [9, 10, 11, 12]]], np.float32)
    out = olr.length_regulate(henc, np.array([[2, 3, 1]]))
    assert out.shape == (1, 6, 4)
    assert np.array_equal(out[0, :, 0], [1, 1, 5, 5, 5, 9])
    # variable totals pad to the longest (15 frames)
    out = olr.length_regulate(np.ones((2, 5, 3), np.float32), np.array([[1, 2, 3, 4, 5], [2, 2, 2, 2, 2]]))
    assert out.shape == (2, 15, 3) and out[1, 10:].sum() == 0


@pytest.mark.gpu
def test_cuda_length_regulator_is_bit_exact():
    import tts_sambert_hifigan_b200 as pkg
    henc, dur, g = _inputs()
    lr = pkg.LengthRegulator()
    out = lr(torch.from_numpy(henc).cuda(), torch.from_numpy(dur).cuda())
    assert out.dtype == torch.float32 and tuple(out.shape) == g["hlr"].shape
    assert np.array_equal(out.cpu().numpy(), g["hlr"])
    # reference KAT (tests/test_length_regulator.py:70-104)
    henc2 = torch.tensor([[[1., 2, 3, 4], [5, 6, 7, 8], [9, 10, 11, 12]]]).cuda()
    out2 = lr(henc2, torch.tensor([[2, 3, 1]]).cuda())
    assert out2.shape == (1, 6, 4) and out2[0, :, 0].tolist() == [1, 1, 5, 5, 5, 9]
    # long sequences cross the 256-thread scan blocks
    h3 = synth.normal(7, (2, 1000, 8))
    d3 = (synth.uniform01(8, 2000).reshape(2, 1000) * 4).astype(np.int64)
    out3 = lr(torch.from_numpy(h3).cuda(), torch.from_numpy(d3).cuda()).cpu().numpy()
    assert np.array_equal(out3, olr.length_regulate(h3, d3))
    with pytest.raises(RuntimeError):
        lr(torch.from_numpy(henc), torch.from_numpy(dur))             # CPU tensors: no fallback


@pytest.mark.gpu
def test_cuda_duration_rounding_matches_reference():
    import tts_sambert_hifigan_b200 as pkg
    g = load_golden("length_regulator")
    log_dur = synth.uniform(203, (B, 400), 3.0)
    d = pkg.durations_from_log(torch.from_numpy(log_dur).cuda()).cpu().numpy()
    assert d.dtype == np.int64 and d.min() >= 1
    # the device evaluates exp in fp64 and rounds to fp32 once: bit-exact against the oracle's definition and,
    # on this fixture, against the golden from the live reference (0 of 1600 differ)
    assert np.array_equal(d, olr.durations_from_log(log_dur))
    assert np.array_equal(d, g["dur_from_log"])
    # on-tie fixture: exact against the oracle; against ATen's CPU expf only near-ties may differ
    ties = np.log(np.arange(1, 4001, dtype=np.float64) + 0.5).astype(np.float32)
    dt = pkg.durations_from_log(torch.from_numpy(ties).cuda()).cpu().numpy()
    assert np.array_equal(dt, olr.durations_from_log(ties))
