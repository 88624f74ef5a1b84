"""ORACLE -- TEST INFRASTRUCTURE ONLY.  numpy restatement of the reference length regulator
(reference models/variance_adaptor.py:171-269) and duration rounding (:746-748).  Pinned against the
live reference by tests/golden/make_golden_lr.py and the reference's own known-answer test
(tests/test_length_regulator.py:70-104)."""
import numpy as np


def durations_from_log(log_dur: np.ndarray) -> np.ndarray:
    # torch.exp(x).round().long() then clamp(min=1); torch.round is half-to-even == np.rint.  exp is taken in
    # float64 and rounded to float32 once (the correctly rounded float32 exponential the reference's float32
    # tensor approximates to within 1 ulp); tests/test_length_regulator.py counts where the live reference's
    # CPU vector expf lands on the other side of a tie.
    d = np.rint(np.exp(log_dur.astype(np.float64)).astype(np.float32)).astype(np.int64)
    return np.maximum(d, 1)


def length_regulate(henc: np.ndarray, dur: np.ndarray) -> np.ndarray:
    dur = np.clip(dur.astype(np.int64), 0, None)                       # :212-219
    outs = [np.repeat(henc[b], dur[b], axis=0) for b in range(henc.shape[0])]   # :232
    tmax = max(o.shape[0] for o in outs)
    out = np.zeros((henc.shape[0], tmax, henc.shape[2]), dtype=henc.dtype)      # :240-264
    for b, o in enumerate(outs):
        out[b, : o.shape[0]] = o
    return out
