"""The C-ABI library loads and exports every symbol include/*.h declares; struct layouts match."""
import ctypes as C
import os
import re

import rabbitvar_b200 as rv

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    names = set()
    for h in ("rabbitvar_b200.h", "rabbitvar_b200_host.h"):
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b(rvh?_[a-z0-9_]+)\s*\(", src))
    return names


def test_every_declared_symbol_is_exported(built):
    L = rv.lib()
    declared = _declared()
    assert len(declared) >= 30
    missing = [n for n in sorted(declared) if not hasattr(L, n)]
    assert not missing, missing
    assert declared == set(rv.ABI_SYMBOLS)


def test_abi_version_and_struct_sizes(built):
    L = rv.lib()
    assert L.rv_abi_version() == 4
    assert C.sizeof(rv.Read) == 32          # fixed per-read header (SURVEY §8d accounting)
    assert C.sizeof(rv.Event) == 160     # 48 fixed bytes + RV_EVENT_KEY_MAX
    assert C.sizeof(rv.Region) == 40
    assert C.sizeof(rv.Variant) == 152


def test_defaults_mirror_reference_cli(built):
    p = rv.default_params()
    # reference defaults: src/Launcher.cpp:299-366
    assert (p.goodq, p.freq, p.lofreq, p.qratio) == (22.5, 0.01, 0.05, 1.5)
    assert (p.vext, p.mismatch, p.minr, p.min_bias_reads, p.read_pos_filter) == (2, 8, 2, 2, 5)
    assert p.samfilter == 0x504 and p.local_realign == 1 and p.bias == 0.05


def test_no_cpu_fallback_without_gpu(built):
    L = rv.lib()
    if L.rv_device_count() > 0:
        return
    h = C.c_void_p()
    p, l = rv.default_params(), rv.default_limits()
    assert L.rv_create(C.byref(h), 0, C.byref(p), C.byref(l)) == -2  # RV_ERR_CUDA
    try:
        rv.Context(0)
    except rv.RabbitVarError as e:
        assert "no CPU fallback" in str(e)
    else:
        raise AssertionError("Context() must fail without a GPU")
