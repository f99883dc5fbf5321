"""The C-ABI library builds, loads on a CPU-only box and exports every symbol include/i2v_b200.h
declares (no compute call is made here)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    """Every function any header under include/ declares (i2v_b200.h = the product ABI, i2v_b200_debug.h = the
    measurement hooks the tools use)."""
    names = set()
    for fn in sorted(os.listdir(os.path.join(ROOT, "include"))):
        if not fn.endswith(".h"):
            continue
        text = open(os.path.join(ROOT, "include", fn)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names.update(re.findall(r"\b(i2v_[a-z0-9_]+)\s*\(", text))
    return sorted(names)


def test_product_header_has_no_debug_hooks():
    text = open(os.path.join(ROOT, "include", "i2v_b200.h")).read()
    assert "i2v_mma_probe" not in text and "i2v_conv_tc_set_trace" not in text


@pytest.fixture(scope="module")
def lib():
    from i2v_b200 import build, capi
    build.build()
    return capi.load()


def test_header_functions_are_exported(lib):
    names = _declared()
    assert len(names) >= 15
    out = subprocess.check_output(["nm", "-D", "--defined-only", lib._name]).decode()
    exported = set(re.findall(r" T (i2v_\w+)", out))
    missing = [n for n in names if n not in exported]
    assert not missing, "declared in the header but not exported: %s" % missing


def test_binding_covers_header(lib):
    from i2v_b200 import capi
    assert sorted(capi.SIGNATURES) == _declared()
    assert capi.version() == 100


def test_no_cpu_fallback():
    """CPU tensors are refused: the product path has no CPU implementation."""
    import torch
    from i2v_b200 import capi
    with pytest.raises(capi.I2VError):
        capi.denorm(torch.zeros(1, 3, 4, 4), torch.zeros(1, 3, 4, 4), 16)


def test_product_never_imports_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's baseline legs may touch oracle/."""
    pkg = os.path.join(ROOT, "image-to-video-i2v-attack_b200")
    files = [os.path.join(pkg, f) for f in os.listdir(pkg) if f.endswith(".py")]
    files += [os.path.join(ROOT, f) for f in ("image_attacks.py", "TPAMI_attack.py", "base_attacks.py", "utils.py", "i2v_b200.py")]
    for f in files:
        src = open(f).read()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
        assert "/root/reference" not in src, f
    for f in os.listdir(os.path.join(pkg, "csrc")):
        src = open(os.path.join(pkg, "csrc", f)).read()
        assert not re.search(r"#include\s*[\"<][^\">]*oracle", src), f


def _sass(obj):
    import shutil
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.isfile(cuobjdump):
        pytest.skip("cuobjdump not available")
    return subprocess.check_output([cuobjdump, "-sass", obj]).decode(errors="replace")


def test_tensor_core_kernels_are_tcgen05_tma_without_waterfall_loops(lib):
    """SASS of the tensor-core convolution object (sm_100a, cross-compiled here): the contraction is tcgen05 (UTCHMMA, also
    the cta_group::2 form), operands and results move by TMA (UTMALDG incl. im2col mode, UTMASTG), accumulators are read
    from tensor memory (LDTM) — and no uniform-datapath instruction sits in a compiler-generated waterfall loop
    (ELECT + R2UR.BROADCAST + BRA.U.ANY): that is what `if (lane == 0)` role selection produced, ~200 cycles per tcgen05.mma
    on the issuing thread; the roles are selected with elect.sync instead."""
    obj = os.path.join(ROOT, "image-to-video-i2v-attack_b200", "build", "conv_tc.o")
    assert os.path.isfile(obj)
    sass = _sass(obj)
    count = lambda pat: len(re.findall(pat, sass))
    assert count(r"\bUTCHMMA\b") > 100
    assert count(r"UTCHMMA\.2CTA") > 0
    assert count(r"\bUTMALDG\.") > 50 and count(r"UTMALDG\.4D\.IM2COL") > 0
    assert count(r"\bUTMASTG\.") > 10
    assert count(r"\bLDTM\b") + count(r"\bLDTM\.") > 50
    assert count(r"BRA\.U\.ANY") == 0, "a waterfall loop is back around a uniform-datapath instruction"
    # the mainloops of the kernels the attack step runs have no emulated integer division per k-step: the only I2F.RP /
    # MUFU.RCP sequences left are per tile (tile -> image / row coordinates)
    for kern in ("conv3x3_halo_kernelILi64ELb1", ):
        body = re.search(r"Function : \S*%s.*?(?=Function : |\Z)" % kern, sass, flags=re.S)
        assert body is not None, kern
        assert len(re.findall(r"BRA\.U\.ANY", body.group(0))) == 0
