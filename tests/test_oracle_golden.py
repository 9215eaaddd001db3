"""CPU: the oracle (oracle/) replays the reference-generated golden vectors bit for bit.

This is what pins the oracle: tests/golden/*.npz were produced by the real reference package
(tests/golden/make_golden.py), and test_oracle_vs_cv2.py checks the same functions against the
installed cv2 live."""
import numpy as np
import pytest

from golden_cases import CASES, E5, golden_blobs, load
from oracle import blobs as oblobs
from oracle import cport
from pygofpid_b200.synth import SyntheticClip


def test_ellipse5_matches_cv2_structuring_element():
    cv = pytest.importorskip("cv2")
    assert np.array_equal(E5, cv.getStructuringElement(cv.MORPH_ELLIPSE, (5, 5)))


def _gray(X):
    Xi = X.astype(np.uint32)
    return ((Xi[..., 0] * 3735 + Xi[..., 1] * 19235 + Xi[..., 2] * 9798 + 16384) >> 15).astype(np.uint8)


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_replays_reference(name):
    case, g = CASES[name], load(name)
    H, W, C = (int(v) for v in g["shape"])
    clip = SyntheticClip(H, W, C, **case["clip"])
    Cd = 1 if case["convert"] else C
    det = {"FD": lambda: cport.FrameDifferencing(5), "MOG2": lambda: cport.MOG2(H, W, Cd),
           "KNN": lambda: cport.KNN(H, W, Cd, int(g["rng_seed"]))}[case["det"]]()
    for t in range(int(g["n"])):
        f = clip.next()
        if case["convert"]:
            f = _gray(f)
        b = case["blur"]
        x = cport.gaussian_blur(f, b[1], b[2]) if b[0] == "gaussian" else cport.box_blur(f, b[1])
        m = det.apply(x.reshape(f.shape))
        for op, se in case["morph"]:
            m = (cport.erode if op == "erode" else cport.dilate)(m, se)
        assert np.array_equal(m, g["mask_%03d" % t]), (name, t)
        want = golden_blobs(g, t)
        got = oblobs.blobs_from_mask(m, area_min=100)
        assert len(got) == len(want), (name, t)
        for a, w in zip(got, want):
            assert np.array_equal(a["contour"], w["contour"])
            assert np.array_equal(a["center"], w["center"]) and np.array_equal(a["bottom"], w["bottom"])


def test_oracle_framediff_matches_reference_class():
    g = load("framediff")
    for tag in ("g", "g1", "c"):
        fd = cport.FrameDifferencing(7.5)
        for f, want in zip(g["frames_" + tag], g["masks_" + tag]):
            assert np.array_equal(fd.apply(f), want.reshape(want.shape[:2]))
