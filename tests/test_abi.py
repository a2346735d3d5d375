"""CPU: the C-ABI library loads and exports every symbol include/harcgpu.h declares; parameter defaults follow
harc:52-63; without a GPU the compute entry points fail loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest

import harness as H


def declared_symbols():
    txt = open(os.path.join(H.ROOT, "include", "harcgpu.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(harcgpu_[a-z_0-9]+)\s*\(", txt)))


def test_every_declared_symbol_is_exported():
    import harc_b200
    lib = harc_b200.load_library()
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(harc_b200.EXPORTS) == names


def test_default_params_follow_harc_script():
    import harc_b200
    for L, want in {100: (50, 18, 49, 50, 81), 250: (125, 93, 124, 125, 156), 36: (18, 7, 17, 18, 28)}.items():
        p = harc_b200.default_params(L)
        assert (p.maxmatch, p.dict_start[0], p.dict_end[0], p.dict_start[1], p.dict_end[1]) == want
        assert (p.thresh, p.thresh_s, p.numdict, p.maxsearch) == (4, 24, 2, 1000)
        o = H.oparams(L)
        assert (o.maxmatch, o.dict_start[0], o.dict_end[1]) == (p.maxmatch, p.dict_start[0], p.dict_end[1])
    with pytest.raises(harc_b200.HarcError):
        harc_b200.default_params(256)  # u8 positions: readlen 256 cannot round-trip


def test_no_cpu_fallback():
    import harc_b200
    lib = harc_b200.load_library()
    if lib.harcgpu_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(harc_b200.HarcError, match="no CUDA device"):
        harc_b200.HarcGpu(100)


def test_product_does_not_link_the_oracle():
    out = os.popen("ldd %s" % os.path.join(H.ROOT, "harc_b200", "libharcgpu.so")).read()
    assert "oracle" not in out
    for root, _, files in os.walk(os.path.join(H.ROOT, "harc_b200")):
        for f in files:
            if f.endswith((".cu", ".cuh", ".h", ".cpp", ".py")):
                assert "oracle" not in open(os.path.join(root, f)).read().lower().replace("# oracle", ""), f


def test_fastq_readlen_is_the_second_line():
    """harc:44 (host side, no GPU needed)."""
    import harc_b200
    assert harc_b200.fastq_readlen(b"@id\nACGTACGTAC\n+\nIIIIIIIIII\n") == 10
    assert harc_b200.fastq_readlen(b"@id\nACGT") == 4
    assert harc_b200.fastq_readlen(b"@id") == -1
