"""CPU: every restatement in oracle/ against the installed OpenCV, bit for bit.

The reference delegates all hot-path arithmetic to cv2 (pipeline.py:463,470,475; helpers.py:139,
156,165; segmentation.py:58-64), so cv2 IS the reference arithmetic.  Skipped when cv2 is absent."""
import numpy as np
import pytest

from oracle import blobs as oblobs
from oracle import cport
from pygofpid_b200.synth import SyntheticClip

cv = pytest.importorskip("cv2")

SHAPES = [(97, 131), (64, 50), (33, 17, 3), (5, 3), (1, 7), (2, 2, 3)]


@pytest.mark.parametrize("shape", SHAPES)
def test_gaussian_and_box(shape):
    rs = np.random.RandomState(0)
    X = rs.randint(0, 256, shape).astype(np.uint8)
    for k, s in [(3, 4), (3, 1), (3, 0), (5, 4), (5, 0), (7, 4), (9, 0), (15, 4), (1, 4), (31, 6)]:
        assert np.array_equal(cv.GaussianBlur(X, (k, k), s), cport.gaussian_blur(X, (k, k), s))
    # the reference's positional call: fun(X, ksize, borderType) -> sigmaX = 4.0
    assert np.array_equal(cv.GaussianBlur(X, (3, 3), cv.BORDER_DEFAULT), cport.gaussian_blur(X, (3, 3), 4.0))
    for k in [(3, 3), (5, 5), (2, 2), (4, 4), (3, 5), (7, 7), (17, 17), (16, 16), (2, 3)]:
        assert np.array_equal(cv.blur(X, k), cport.box_blur(X, k))
        assert np.array_equal(cv.blur(X, k, cv.BORDER_REPLICATE), cport.box_blur(X, k))  # 3rd arg = dst


def test_gaussian_taps():
    for n, s, want in [(3, 4, [84, 88, 84]), (3, 1, [70, 116, 70]), (5, 4, [48, 53, 54, 53, 48]),
                       (7, 4, [31, 36, 40, 42, 40, 36, 31]), (3, 0, [64, 128, 64])]:
        assert list(cport.gauss_kernel_q8(n, s)) == want


def test_morphology():
    rs = np.random.RandomState(1)
    ses = [cv.getStructuringElement(cv.MORPH_RECT, (5, 5)), cv.getStructuringElement(cv.MORPH_ELLIPSE, (15, 15)),
           cv.getStructuringElement(cv.MORPH_CROSS, (5, 5)), np.ones((2, 4), np.uint8),
           cv.getStructuringElement(cv.MORPH_ELLIPSE, (7, 5)), np.array([[1, 0, 0], [1, 0, 0], [1, 1, 5]], np.uint8)]
    for shape in [(40, 37), (5, 3), (1, 9)]:
        M = (rs.randint(0, 3, shape) * 127.5).astype(np.uint8)
        for se in ses:
            assert np.array_equal(cv.erode(M, se), cport.erode(M, se))
            assert np.array_equal(cv.dilate(M, se), cport.dilate(M, se))


@pytest.mark.parametrize("C,sigma", [(1, 6), (3, 6), (1, 25), (3, 25)])
def test_mog2(C, sigma):
    H, W = 40, 56
    m, o = cv.createBackgroundSubtractorMOG2(), cport.MOG2(H, W, C)
    clip = SyntheticClip(H, W, C, seed=10 * C + sigma, n_obj=4, noise_sigma=sigma)
    for t in range(260):
        f = clip.next()
        assert np.array_equal(m.apply(f), o.apply(f)), t
    assert np.array_equal(m.getBackgroundImage(), o.background())


@pytest.mark.parametrize("C", [1, 3])
def test_knn(C):
    H, W = 37, 53
    cv.setRNGSeed(12345)
    m, o = cv.createBackgroundSubtractorKNN(), cport.KNN(H, W, C, 12345)
    clip = SyntheticClip(H, W, C, seed=C, n_obj=4)
    for t in range(160):
        f = clip.next()
        assert np.array_equal(m.apply(f), o.apply(f)), t


def test_external_contours_area_moments():
    rs = np.random.RandomState(5)
    total = 0
    for it in range(30):
        H, W = rs.randint(20, 80), rs.randint(20, 80)
        if it % 3 == 0:
            M = (rs.rand(H, W) < 0.5).astype(np.uint8) * 255
        else:
            M = np.zeros((H, W), np.uint8)
            for _ in range(8):
                x, y = rs.randint(0, W), rs.randint(0, H)
                cv.rectangle(M, (x, y), (x + rs.randint(2, 30), y + rs.randint(2, 30)),
                             int(rs.choice([127, 255])), int(rs.choice([-1, 1, 2])))
            M[rs.rand(H, W) < 0.05] = 0
        a = cv.findContours(M, cv.RETR_EXTERNAL, cv.CHAIN_APPROX_NONE)[0]
        b = oblobs.find_external_contours(M)
        assert len(a) == len(b)
        for ca, cb in zip(a, b):
            assert np.array_equal(ca, cb)
            pts = [tuple(p) for p in cb.reshape(-1, 2).tolist()]
            assert oblobs.contour_area(pts) == cv.contourArea(ca)
            mm = cv.moments(ca)
            assert (mm['m00'], mm['m10'], mm['m01']) == oblobs.contour_moments(pts)
            total += 1
    assert total > 100
