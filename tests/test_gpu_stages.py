"""GPU parity of the stage-level C-ABI entry points against the CPU oracle (bit-exact).

gofpid_blur_u8 <-> cv.GaussianBlur / cv.blur   (reference call site pipeline.py:463)
gofpid_morph_u8 <-> cv.erode / cv.dilate chain (pipeline.py:473-475)
gofpid_ccl_u8   <-> canonical 8-connected labels (device half of pipeline.py:555-570)
"""
import numpy as np
import pytest

from oracle import cport

pytestmark = pytest.mark.gpu

SHAPES = [(97, 131), (64, 50), (33, 17, 3), (5, 3), (1, 7), (2, 2, 3), (240, 320), (20, 32, 3),
          (64, 32), (300, 517, 3), (130, 260)]


def _torch():
    import torch
    return torch


def ellipse(h, w):
    """cv.getStructuringElement(MORPH_ELLIPSE) of OpenCV 4.13 restated (rows of an inscribed ellipse)."""
    import cv2
    return cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (w, h))


@pytest.mark.parametrize("shape", SHAPES)
def test_gaussian_blur_bitexact(shape):
    torch = _torch()
    from pygofpid_b200.hotpath import blur_u8
    rs = np.random.RandomState(hash(shape) % 1000)
    X = rs.randint(0, 256, shape).astype(np.uint8)
    d = torch.from_numpy(X).cuda()
    for k, s in [(3, 4.0), (3, 1.0), (3, 0.0), (5, 4.0), (5, 0.0), (7, 4.0), (9, 0.0), (15, 4.0),
                 (1, 4.0), (31, 6.0)]:
        want = cport.gaussian_blur(X, (k, k), s).reshape(X.shape)
        got = blur_u8(d, "gaussian", (k, k), s).cpu().numpy()
        assert np.array_equal(got, want), (shape, k, s, int((got != want).sum()))
    # non-square kernel
    want = cport.gaussian_blur(X, (5, 3), 2.0).reshape(X.shape)
    got = blur_u8(d, "gaussian", (5, 3), 2.0).cpu().numpy()
    assert np.array_equal(got, want)


@pytest.mark.parametrize("shape", SHAPES)
def test_box_blur_bitexact(shape):
    torch = _torch()
    from pygofpid_b200.hotpath import blur_u8
    rs = np.random.RandomState(7)
    X = rs.randint(0, 256, shape).astype(np.uint8)
    d = torch.from_numpy(X).cuda()
    for k in [(3, 3), (5, 5), (2, 2), (4, 4), (3, 5), (7, 7), (17, 17), (16, 16), (2, 3), (1, 1)]:
        want = cport.box_blur(X, k).reshape(X.shape)
        got = blur_u8(d, "box", k).cpu().numpy()
        assert np.array_equal(got, want), (shape, k, int((got != want).sum()))


def test_blur_none_is_identity():
    torch = _torch()
    from pygofpid_b200.hotpath import blur_u8
    X = np.random.RandomState(1).randint(0, 256, (37, 53, 3)).astype(np.uint8)
    got = blur_u8(torch.from_numpy(X).cuda(), "none", (1, 1)).cpu().numpy()
    assert np.array_equal(got, X)


def _grey_mask(rs, shape):
    return (rs.randint(0, 3, shape) * 127.5).astype(np.uint8)


MORPH_SHAPES = [(40, 37), (5, 3), (1, 9), (97, 131), (64, 50), (200, 300), (33, 260)]


@pytest.mark.parametrize("shape", MORPH_SHAPES)
def test_morph_single_ops_bitexact(shape):
    torch = _torch()
    from pygofpid_b200.hotpath import morph_u8
    rs = np.random.RandomState(3)
    M = _grey_mask(rs, shape)
    M2 = rs.randint(0, 256, shape).astype(np.uint8)
    ses = [np.ones((3, 3), np.uint8), np.ones((5, 5), np.uint8), ellipse(15, 15), ellipse(5, 7),
           np.ones((2, 4), np.uint8), np.ones((1, 1), np.uint8),
           np.array([[0, 1, 0], [1, 1, 1], [0, 1, 0]], np.uint8),
           np.array([[1, 0, 0], [1, 0, 0], [1, 1, 5]], np.uint8),
           np.array([[1, 0, 1, 0, 1]], np.uint8), np.ones((7, 1), np.uint8),
           np.ones((4, 6), np.uint8)]
    for img in (M, M2):
        d = torch.from_numpy(img).cuda()
        for se in ses:
            for name, fn in (("erode", cport.erode), ("dilate", cport.dilate)):
                want = fn(img, se)
                got = morph_u8(d, [(name, se)]).cpu().numpy()
                assert np.array_equal(got, want), (shape, se.shape, name, int((got != want).sum()))


@pytest.mark.parametrize("shape", MORPH_SHAPES)
def test_morph_chains_bitexact(shape):
    torch = _torch()
    from pygofpid_b200.hotpath import morph_u8
    rs = np.random.RandomState(4)
    M = _grey_mask(rs, shape)
    r5 = np.ones((5, 5), np.uint8)
    r3 = np.ones((3, 3), np.uint8)
    chains = [
        [("erode", r5), ("dilate", r5)],                                   # GOFPID default: open
        [("erode", r5), ("dilate", r5), ("dilate", r5), ("erode", r5)],   # open then close
        [("erode", ellipse(15, 15)), ("dilate", ellipse(15, 15))],        # BASELINE config 3
        [("dilate", r3), ("erode", ellipse(7, 5)), ("dilate", np.ones((2, 4), np.uint8))],
        [],
    ]
    d = torch.from_numpy(M).cuda()
    for chain in chains:
        want = M
        for name, se in chain:
            want = (cport.erode if name == "erode" else cport.dilate)(want, se)
        got = morph_u8(d, chain).cpu().numpy()
        assert np.array_equal(got, want), (shape, [(n, s.shape) for n, s in chain],
                                           int((got != want).sum()))


BIT_SHAPES = MORPH_SHAPES + [(70, 600), (300, 257), (65, 1000)]


@pytest.mark.parametrize("shape", BIT_SHAPES)
def test_morph_bitplane_path_bitexact(shape):
    """The pipeline's bit-plane kernel (threshold decomposition) against the oracle, for every
    detector alphabet: {255}, {127, 255}, {84, 85}, three levels, W not a multiple of 16/32/256."""
    torch = _torch()
    from pygofpid_b200.hotpath import morph_u8
    rs = np.random.RandomState(11)
    r5 = np.ones((5, 5), np.uint8)
    r3 = np.ones((3, 3), np.uint8)
    lshape = np.array([[1, 0, 0], [1, 0, 0], [1, 1, 5]], np.uint8)
    noanchor = np.array([[1, 0, 0, 0, 1]], np.uint8)  # anchor element is 0
    chains = [
        [("erode", r5), ("dilate", r5)],
        [("erode", r5), ("dilate", r5), ("dilate", r5), ("erode", r5)],
        [("erode", ellipse(15, 15)), ("dilate", ellipse(15, 15))],
        [("dilate", r3), ("erode", ellipse(7, 5)), ("dilate", np.ones((2, 4), np.uint8))],
        [("erode", lshape)], [("dilate", lshape)], [("erode", np.ones((4, 6), np.uint8))],
        [("dilate", np.ones((7, 1), np.uint8)), ("erode", np.ones((1, 9), np.uint8))],
        [("erode", noanchor)], [("dilate", noanchor), ("erode", noanchor)],
        [("dilate", np.ones((31, 31), np.uint8))], [("erode", ellipse(31, 31)), ("dilate", ellipse(31, 31))],
    ]
    H, W = shape
    for levels in ([255], [127, 255], [84, 85], [1, 128, 200]):
        alphabet = np.array([0] + levels, np.uint8)
        masks = [alphabet[rs.randint(0, len(alphabet), shape)],                      # dense noise
                 alphabet[(rs.rand(H, W) < 0.03) * rs.randint(1, len(alphabet), shape)],  # sparse
                 _grey_mask(rs, shape) if levels == [127, 255] else None,
                 np.zeros(shape, np.uint8), np.full(shape, levels[-1], np.uint8)]
        for M in masks:
            if M is None:
                continue
            d = torch.from_numpy(np.ascontiguousarray(M)).cuda()
            for chain in chains:
                want = M
                for name, se in chain:
                    want = (cport.erode if name == "erode" else cport.dilate)(want, se)
                got = morph_u8(d, chain, levels=levels).cpu().numpy()
                assert np.array_equal(got, want), (shape, levels, [(n, s.shape) for n, s in chain],
                                                   int((got != want).sum()))


def _blob_mask(rs, H, W, density=None):
    import cv2
    if density is not None:
        return ((rs.rand(H, W) < density) * rs.choice([127, 255], (H, W))).astype(np.uint8)
    M = np.zeros((H, W), np.uint8)
    for _ in range(12):
        x, y = rs.randint(0, W), rs.randint(0, H)
        w, h = rs.randint(2, max(3, W // 3)), rs.randint(2, max(3, H // 3))
        cv2.rectangle(M, (x, y), (x + w, y + h), int(rs.choice([127, 255])), int(rs.choice([-1, 1, 2])))
    M[rs.rand(H, W) < 0.05] = 0
    return M


@pytest.mark.parametrize("shape", [(20, 20), (16, 128), (17, 129), (40, 300), (97, 131), (240, 320),
                                   (1, 50), (50, 1), (200, 517)])
@pytest.mark.parametrize("path", ["runs", "tiled", "runs_cap64"])
def test_ccl_labels_bitexact(shape, path, monkeypatch):
    """Both labelling paths: run-based in shared memory (the usual one) and the tiled union-find
    fallback (forced, or reached because the frame has more runs than the capacity)."""
    torch = _torch()
    from pygofpid_b200.hotpath import ccl_u8
    if path == "tiled":
        monkeypatch.setenv("GOFPID_CCL_FORCE_TILED", "1")
    elif path == "runs_cap64":
        monkeypatch.setenv("GOFPID_RUNS_CAP", "64")
    rs = np.random.RandomState(11)
    H, W = shape
    masks = [_blob_mask(rs, H, W), _blob_mask(rs, H, W, 0.5), _blob_mask(rs, H, W, 0.62),
             np.zeros(shape, np.uint8), np.full(shape, 255, np.uint8)]
    # a spiral / comb that needs many merge rounds
    comb = np.zeros(shape, np.uint8)
    comb[::2, :] = 255
    comb[1::4, 0] = 255
    comb[3::4, -1] = 255
    masks.append(comb)
    for M in masks:
        want = cport.ccl8(M)
        got = ccl_u8(torch.from_numpy(M).cuda()).cpu().numpy()
        assert np.array_equal(got, want), (shape, int((got != want).sum()))
