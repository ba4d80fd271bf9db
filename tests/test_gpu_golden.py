"""The CUDA path must reproduce the committed golden fixtures (tests/golden): counts, flags and -- in the
deterministic mode -- every statistic bit for bit; the atomic mode within 1e-6."""
import numpy as np
import pytest

from gritas_b200 import cabi
from tests import helpers as H

pytestmark = pytest.mark.gpu


class _Ref:
    def __init__(self, d):
        self.__dict__.update(d)


@pytest.mark.parametrize("mode", H.MODES)
@pytest.mark.parametrize("name", H.GOLDEN)
def test_golden(name, mode):
    cfg, files, flags, norm, raw, l2d, l3d, p1, p2 = H.load_golden(name)
    fin = [H.filter_in(c) for c in files]
    out, gfl = H.gpu_run(cfg, files, cfg.nr, mode, fin)
    assert all(np.array_equal(a, b) for a, b in zip(gfl, flags))
    H.compare(out, _Ref(norm), l2d, l3d, cfg.nr, exact=H.is_det(mode), tol=1e-6)
    rout, _ = H.gpu_run(cfg, files, cfg.nr, mode, fin, norm=False)
    rr = dict(raw)
    H.compare(rout, _Ref(rr), l2d, l3d, cfg.nr, exact=H.is_det(mode), tol=1e-6, raw=True)
