"""Pins oracle/lsap.c (restatement of scipy's rectangular LSAP) against the installed scipy and against
the committed scipy known-answer vectors.  CPU only."""
import os

import numpy as np
import pytest
import torch
from scipy.optimize import linear_sum_assignment

from oracle.lsap import lsap_c


def regen_lsap_cases(golden_dir):
    gold = torch.load(os.path.join(golden_dir, "lsap_scipy.pt"), weights_only=False)
    g = torch.Generator().manual_seed(gold["generator_seed"])
    it = iter(gold["cases"])
    for nr, nc in gold["shapes"]:
        for kind in ("normal", "int", "dupcol"):
            if kind == "normal":
                c = torch.randn(nr, nc, generator=g)
            elif kind == "int":
                c = torch.randint(0, 4, (nr, nc), generator=g).float()
            else:
                c = torch.randn(nr, max(1, (nc + 4) // 5), generator=g).repeat_interleave(5, 1)[:, :nc].contiguous()
            case = next(it)
            assert abs(float(c.double().sum()) - case["seed_cost_sum"]) < 1e-9
            yield c, case


def test_c_oracle_matches_committed_scipy_vectors(golden_dir):
    n = 0
    for c, case in regen_lsap_cases(golden_dir):
        i, j = lsap_c(c.numpy())
        assert np.array_equal(i, case["rows"].numpy()) and np.array_equal(j, case["cols"].numpy()), case["shape"]
        n += 1
    assert n == 33


@pytest.mark.parametrize("seed", range(4))
def test_c_oracle_matches_installed_scipy(seed):
    rng = np.random.default_rng(seed)
    for t in range(120):
        nr, nc = rng.integers(1, 48, 2)
        if t % 3 == 0:
            c = rng.standard_normal((nr, nc)).astype(np.float32)
        elif t % 3 == 1:
            c = rng.integers(0, 3, (nr, nc)).astype(np.float32)
        else:
            c = np.repeat(rng.standard_normal((nr, (nc + 2) // 3)).astype(np.float32), 3, 1)[:, :nc]
        a, b = linear_sum_assignment(c), lsap_c(c)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_empty_and_errors():
    i, j = lsap_c(np.zeros((0, 5), np.float32))
    assert len(i) == 0 and len(j) == 0
    with pytest.raises(ValueError):
        lsap_c(np.array([[np.nan, 1.0]], np.float32))
