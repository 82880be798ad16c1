"""CPU-only: the C-ABI library loads and exports what include/sonarfe.h declares; host-side
logic of the drop-in classes (no compute calls -- there is no GPU here)."""
import json
import os
import re

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(REPO, "include", "sonarfe.h")).read()
    return sorted(set(re.findall(r"SFE_API[^;(]*?\b(sfe_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from sonar_slam_b200 import _lib
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 8
    for n in names:
        assert hasattr(lib, n), f"libsonarfe.so does not export {n}"
    assert lib.sfe_version() == 201


def test_no_cpu_fallback_context_creation_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from sonar_slam_b200 import _lib
    with pytest.raises(_lib.SonarFEError):
        _lib.Context(0)


def test_cfar_class_threshold_factors_match_reference(golden_dir):
    from sonar_slam_b200.bruce_slam.CFAR import CFAR
    for g in json.load(open(os.path.join(golden_dir, "cfar_tau.json"))):
        if "raises" in g:
            with pytest.raises(ValueError, match=g["raises"]):
                CFAR(g["Ntc"], g["Ngc"], g["Pfa"], g["rank"])
            continue
        c = CFAR(g["Ntc"], g["Ngc"], g["Pfa"], g["rank"])
        for alg in ("CA", "SOCA", "GOCA", "OS"):
            got = getattr(c, "threshold_factor_" + alg)
            assert got == pytest.approx(g[alg], rel=1e-12, abs=0), (g, alg)
        assert str(c) == g["str"]
        assert c.params["SOCA"][:2] == (g["Ntc"] // 2, g["Ngc"] // 2)
        assert c.params["OS"][2] == (g["rank"] if g["rank"] is not None else g["Ntc"] / 2)


def test_cfar_class_asserts_like_reference():
    from sonar_slam_b200.bruce_slam.CFAR import CFAR
    for bad in [(41, 10, 0.1, 1), (40, 9, 0.1, 1), (40, 10, 0.1, 40), (40, 10, 0.1, -1)]:
        with pytest.raises(AssertionError):
            CFAR(*bad)


def test_cfar_module_argument_checks_happen_before_any_device_work():
    from sonar_slam_b200.bruce_slam import cfar
    img = np.zeros((64, 32), np.uint8)
    with pytest.raises(TypeError):
        cfar.soca(np.zeros((4, 4, 4)), 20, 5, 1.0)      # not 2-D
    with pytest.raises(TypeError):
        cfar.soca(img, 20.0, 5, 1.0)                    # pybind's `int train_hs` rejects floats
    with pytest.raises(TypeError):
        cfar.os(img, 20, 5, 10.0, 1.0)                  # rank=None -> Ntc/2 float (CFAR.py:24) is rejected too
    with pytest.raises(TypeError):
        cfar.ca(np.zeros((4, 4), dtype=complex), 1, 1, 1.0)


def test_compressed_ping_is_decoded_like_the_reference():
    """feature_extraction.py:210-213: cv2.imdecode(..., IMREAD_COLOR) then BGR -> gray; PNG is lossless, so the decoded
    ping equals the raw one; the uncompressed branch passes arrays through."""
    import types
    cv2 = pytest.importorskip("cv2")
    from sonar_slam_b200 import synth
    from sonar_slam_b200.bruce_slam.feature_extraction import FeatureExtraction
    img = synth.make_frame(seed=1)
    ok, png = cv2.imencode(".png", img)
    assert ok
    fe = FeatureExtraction()
    fe.compressed_images = True
    got = fe.ping_image(types.SimpleNamespace(ping=types.SimpleNamespace(data=png.tobytes())))
    want = cv2.cvtColor(np.array(cv2.imdecode(np.frombuffer(png.tobytes(), np.uint8), cv2.IMREAD_COLOR)).astype(np.uint8),
                        cv2.COLOR_BGR2GRAY)
    assert got.dtype == np.uint8 and np.array_equal(got, want) and np.array_equal(got, img)
    with pytest.raises(ValueError):
        fe.ping_image(types.SimpleNamespace(ping=types.SimpleNamespace(data=b"not an image")))
    fe.compressed_images = False
    assert np.array_equal(fe.ping_image(types.SimpleNamespace(ping=img)), img)
