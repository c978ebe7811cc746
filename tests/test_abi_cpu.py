"""CPU-only checks: the C-ABI library builds (nvcc cross-compile), loads and exports every symbol
include/zeggs_b200.h declares; host-side table construction matches the oracle; no compute calls."""
import os
import re

import numpy as np
import pytest

from tests._util import ROOT, ensure_built


@pytest.fixture(scope="module")
def built():
    ensure_built()
    from zeggs_b200 import _lib
    return _lib


def test_library_exports_every_header_symbol(built):
    hdr = open(os.path.join(ROOT, "include", "zeggs_b200.h")).read()
    declared = set(re.findall(r"\b(zeggs_[a-z0-9_]+)\s*\(", hdr))
    l = built.lib()
    bound = {s[0] for s in built.SYMBOLS}
    assert declared == bound, (declared - bound, bound - declared)
    for name in declared:
        assert getattr(l, name) is not None
    assert l.zeggs_version() >= 100


def test_ctypes_mirrors_have_the_library_struct_layout(built):
    """Every args struct of the header is mirrored in _lib.py: same size as the compiled C struct (a drifted mirror would make the
    library read garbage pointers), and every struct the header defines has a mirror."""
    import ctypes
    hdr = open(os.path.join(ROOT, "include", "zeggs_b200.h")).read()
    defined = set(re.findall(r"^\}\s*(zeggs_[a-z0-9_]+);", hdr, flags=re.M))
    mirrors = built.struct_mirrors()
    assert defined == set(mirrors), (defined - set(mirrors), set(mirrors) - defined)
    l = built.lib()
    for name, cls in mirrors.items():
        assert l.zeggs_struct_size(name.encode()) == ctypes.sizeof(cls), name
    assert l.zeggs_struct_size(b"no_such_struct") == 0


def test_mel_frame_count_rule(built):
    l = built.lib()
    from oracle import mel_oracle
    for ns, hop in [(160000, 200), (160000, 160), (16001, 200), (12345, 160), (800, 200), (500, 200)]:
        assert l.zeggs_mel_num_frames(ns, 800, hop) == mel_oracle.num_frames(ns, 800, hop)


def test_workspace_and_pack_sizes(built):
    l = built.lib()
    assert l.zeggs_decoder_packed_bytes(1024, 64, 64) > 4 * (1024 * 1134 * 4 + 12 * 1024 * 1024)
    assert l.zeggs_decoder_packed_bytes(1000, 64, 64) == 0          # H % 16 != 0 -> unsupported
    a = l.zeggs_decoder_workspace_bytes(32, 256, 1024, 64, 64, 1)
    b = l.zeggs_decoder_workspace_bytes(32, 256, 1024, 64, 64, 0)
    assert a > b > 0


def test_host_filterbank_matches_oracle():
    from oracle import mel_oracle
    from zeggs_b200.audio import mel_filterbank
    a = mel_filterbank(800, 16000, 80, 20, 7600, True)
    b = mel_oracle.mel_filterbank(800, 16000, 80, 20, 7600, True)
    assert np.array_equal(a, b)
    assert int((a != 0).sum()) == 742


def test_compute_path_refuses_cpu_tensors(built):
    import torch
    from zeggs_b200 import modules, _lib
    dec = modules.Decoder(1134, 1131, 64, 64, 64, 2)
    z = torch.zeros
    with pytest.raises(_lib.ZeggsError):
        dec(z(1, 3), z(1, 4), z(1, 3), z(1, 3), z(1, 75, 3), z(1, 75, 2, 3), z(1, 75, 3), z(1, 75, 3), z(1, 4, 3),
            z(1, 4, 64), z(1, 4, 64), None, z(1134), torch.ones(1134), z(1131), torch.ones(1131), 1 / 60)
