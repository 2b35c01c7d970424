"""Shared helpers for the parity tests."""
import importlib.util
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name)))


def hostemu():
    spec = importlib.util.spec_from_file_location("nefii_hostemu", os.path.join(ROOT, "tests", "hostemu", "__init__.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def rel_stats(got, ref, abs_floor=1e-6):
    """(fraction of lanes within rel 1e-4, p99 rel err, max rel err) with an absolute floor."""
    got = got.detach().double().cpu().flatten()
    ref = ref.detach().double().cpu().flatten()
    rel = (got - ref).abs() / (ref.abs() + abs_floor)
    rel = torch.nan_to_num(rel, nan=float("inf"))
    frac = (rel <= 1e-4).double().mean().item()
    k = max(1, int(0.99 * rel.numel()))
    p99 = rel.kthvalue(k)[0].item()
    return frac, p99, rel.max().item()
