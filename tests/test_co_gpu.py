"""GPU parity of the object x object correlation: picca_b200.co (pb2_neigh_* + pb2_co_pairs through
the C ABI) against the live reference's golden vectors (tests/golden/golden_co.npz): num_pairs
bit-exact per HEALPix pixel, weighted sums within 1e-9."""
import os

import numpy as np
import pytest

from tests.golden import cases
from tests.test_oracle_golden import setup_co

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_co.npz")


@pytest.mark.parametrize("name", sorted(cases.CO_CASES))
def test_co_matches_reference_golden(name):
    from picca_b200 import co
    gold = np.load(GOLD)["co_%s" % name]
    objs = setup_co(co, cases.CO_CASES[name])
    rows = []
    for hp in sorted(objs):
        co.fill_neighs([hp])
        res = co.compute_xi([hp])
        assert len(res) == 5 and res[4].dtype == np.int64
        rows.append(np.stack([np.asarray(r, dtype=np.float64) for r in res[:4]] +
                             [np.asarray(res[4], dtype=np.int64).view(np.float64)]))
    rows = np.stack(rows)
    assert np.array_equal(rows[:, 4].view(np.int64), gold[:, 4].view(np.int64))
    assert rows[:, 4].view(np.int64).sum() > 500
    for k in range(4):
        np.testing.assert_allclose(rows[:, k], gold[:, k], rtol=1e-9, atol=1e-300)
