"""The C-ABI library builds, loads and exports every entry point include/afv.h declares (no GPU needed),
and fails loudly (no CPU fallback) when no CUDA device is usable."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "afv.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(afv_[a-z_0-9]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    pkg = g._load_pkg()
    lib = pkg.lib()
    names = _declared()
    assert len(names) >= 17
    for n in names:
        assert hasattr(lib, n), "missing export %s" % n
    assert b"sm_100a" in lib.afv_version()


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import __graft_entry__ as g
    pkg = g._load_pkg()
    with pytest.raises(pkg.AfvError) as e:
        pkg.FeatureExtractor("orb32")
    assert "no CPU fallback" in str(e.value)


def test_unsupported_feature_is_an_error_not_a_fallback():
    import __graft_entry__ as g
    pkg = g._load_pkg()
    h = C.c_void_p()
    # surf64 (feature id 3): no extractor in this library -> an error code, never a substitute
    rc = pkg.lib().afv_extractor_create(C.byref(h), 3, 1000, 8, C.c_float(1.5), C.c_float(34.0), 0, 1, 640, 480)
    assert rc == -5 and not h.value
    import torch
    if not torch.cuda.is_available():                       # sift128 / akaze61 / brisk48 are built, but never on the CPU
        for fid, sf, th in ((5, 2.0, 10.0), (1, 1.1892, 5e-4), (2, 1.5, 34.0)):
            rc = pkg.lib().afv_extractor_create(C.byref(h), fid, 1000, 8, C.c_float(sf), C.c_float(th), 0, 1, 640, 480)
            assert rc == -2 and not h.value


def test_cpp_host_mirror_builds_and_links():
    """The C++ classes that keep the reference's FeatureExtractor / FeatureMatcher signatures compile and link
    against the C-ABI library without OpenCV."""
    import subprocess
    import __graft_entry__ as g
    g.build()
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "anyfeature-vslam_b200", "host")])
    assert os.path.exists(os.path.join(ROOT, "anyfeature-vslam_b200", "host", "host_api_test"))
