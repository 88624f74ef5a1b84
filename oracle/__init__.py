"""ORACLE -- test infrastructure only (see hifigan_oracle.c / torch_port.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this package.  The product package never does.
"""
from .c_oracle import build_c_oracle, forward_c            # noqa: F401
from .torch_port import forward_torch, fold_weight_norm    # noqa: F401
from .split_plan_model import forward_split_plan, forward_mode_model   # noqa: F401
