"""GPU: the CUDA path against the committed golden vectors of the reference's own binaries (tests/golden)."""
import os

import numpy as np
import pytest

from test_oracle_golden import FIX, GOLD, compare_dir, fixture_L, parse_dump, stage_dir

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", FIX)
def test_gpu_matches_reference_golden(name, tmp_path):
    import harc_b200
    L = fixture_L(name)
    d = stage_dir(str(tmp_path), name, ["in"])
    ctx = harc_b200.HarcGpu(L, walkers=1, file_sets=1)
    ctx.reorder_dir(d)
    gold = parse_dump(os.path.join(GOLD, name, "dict1.bin"))
    for l in range(2):
        k, c, i = ctx.dump_dict(1, l)
        assert np.array_equal(k, gold[l][0]) and np.array_equal(c, gold[l][1]) and np.array_equal(i, gold[l][2])
    assert compare_dir(d, name, "s1") == []
    ctx.encode_dir(d)
    gold2 = parse_dump(os.path.join(GOLD, name, "dict2.bin"))
    for l in range(2):
        k, c, i = ctx.dump_dict(2, l)
        assert np.array_equal(k, gold2[l][0]) and np.array_equal(c, gold2[l][1]) and np.array_equal(i, gold2[l][2])
    ctx.close()
    assert compare_dir(d, name, "s2") == []
