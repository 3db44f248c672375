"""Image.toColors / toBitmap / saveBitmap (Image.fs:37-90).
The reference's bitmap is not reproducible (shared System.Random inside a parallel map), so the dither is
specified here as Philox4x32-10(seed, pixel index); parity is against the oracle's restatement with the same
dither ("unpinned", like everything else)."""
import struct

import numpy as np
import pytest

import fraytracer_b200 as ft
from fraytracer_b200 import scenes
from oracle import oracle as O


def test_philox_known_answers():
    # Random123 known-answer vectors for philox4x32-10
    assert [hex(v) for v in O.philox4x32_10([0, 0, 0, 0], [0, 0])] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    assert [hex(v) for v in O.philox4x32_10([0xffffffff] * 4, [0xffffffff] * 2)] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]


def test_oracle_to_bitmap_by_hand():
    # gamma 1 (pow(x,1) == x), 2 x 1 image whose maximum is 2.0: bytes = round(c/2 * 254.5 + u), u in [0,1)
    img = np.array([[[2.0, 1.0, 0.0]], [[0.5, 0.0, 2.0]]], np.float32)  # [X=2, Y=1, 3]
    bmp, mx = O.to_bitmap(img, 1.0, seed=7)
    assert mx == 2.0 and bmp.shape == (1, 2, 3)
    # scanline 0: column 0 = image[X-1-0 = 1, 0], column 1 = image[0, 0]; bytes are B,G,R
    for col, x in ((0, 1), (1, 0)):
        for ch, c in zip((2, 1, 0), img[x, 0]):
            lo = c / 2 * 254.5
            assert np.floor(lo) <= bmp[0, col, ch] <= np.ceil(lo + 1) and bmp[0, col, ch] <= 255
    assert bmp[0, 1, 2] in (254, 255) and bmp[0, 1, 0] in (0, 1)
    # the maximum is clamped from below: an all-dark image is divided by 0.01, not by its tiny maximum
    dark = np.full((2, 2, 3), 0.001, np.float32)
    bmp, mx = O.to_bitmap(dark, 1.0, seed=1)
    assert abs(mx - 0.001) < 1e-9 and set(np.unique(bmp)) <= {25, 26}  # 0.1 * 254.5 = 25.45 (+ dither)


def test_save_bitmap_roundtrip(tmp_path):
    rng = np.random.default_rng(0)
    for X, Y in ((5, 3), (8, 2), (1, 1)):
        bmp = rng.integers(0, 256, size=(Y, X, 3), dtype=np.uint8)
        path = tmp_path / f"t{X}x{Y}.bmp"
        ft.Image.saveBitmap(str(path), bmp)
        raw = path.read_bytes()
        assert raw[:2] == b"BM"
        size, off = struct.unpack_from("<I4xI", raw, 2)
        w, h, planes, bpp = struct.unpack_from("<iiHH", raw, 18)
        assert (size, off, w, h, planes, bpp) == (len(raw), 54, X, Y, 1, 24)
        row = (3 * X + 3) & ~3
        got = np.stack([np.frombuffer(raw, np.uint8, 3 * X, off + (Y - 1 - r) * row).reshape(X, 3) for r in range(Y)])
        assert np.array_equal(got, bmp)  # BMP rows are stored bottom-up


@pytest.mark.gpu
def test_gpu_to_bitmap_matches_oracle():
    w = scenes.default_scene(size=300)
    tracer = ft.SceneTracer(w.scene, device=0)
    img, _, _ = tracer.render(w.camera, w.size, w.epsilon, w.length)
    got = ft.Image.toBitmap(2.2, 19, img)
    want, _ = O.to_bitmap(img, 2.2, 19)
    diff = np.abs(got.astype(int) - want.astype(int))
    # MathF.Pow: glibc powf (oracle) vs CUDA powf differ by a few ulp; a byte moves by one count only when
    # c*254.5 + dither lands within ~1e-5 of a rounding boundary
    assert diff.max() <= 1 and (diff != 0).mean() < 1e-4
    assert got.shape == (300, 300, 3) and got.max() >= 254
    # gamma 1: pow(x, 1) == x exactly in both libraries: bytes identical
    assert np.array_equal(ft.Image.toBitmap(1.0, 3, img), O.to_bitmap(img, 1.0, 3)[0])
    # ragged size and another seed
    small = np.ascontiguousarray(img[130:167, 120:173])  # 37 x 53 crop from the middle of the frame
    assert np.abs(ft.Image.toBitmap(2.2, 5, small).astype(int) - O.to_bitmap(small, 2.2, 5)[0].astype(int)).max() <= 1
    assert not np.array_equal(ft.Image.toBitmap(2.2, 5, small), ft.Image.toBitmap(2.2, 6, small))


@pytest.mark.gpu
def test_gpu_image_max_and_device_tonemap_with_given_max():
    import ctypes as C

    import torch

    from fraytracer_b200 import _abi as A

    lib = ft.load_library()
    rng = np.random.default_rng(4)
    img = rng.uniform(0, 3, size=(70, 45, 3)).astype(np.float32)
    img[17, 30, 1] = 9.5
    d = torch.from_numpy(img).cuda()
    d_max = torch.zeros(1, dtype=torch.float32, device="cuda")
    A.check(lib.ft_image_max_device(C.c_void_p(d.data_ptr()), img.size, C.c_void_p(d_max.data_ptr()), None))
    assert float(d_max.item()) == 9.5
    out = torch.zeros((45, 70, 3), dtype=torch.uint8, device="cuda")
    A.check(lib.ft_tonemap_device(C.c_void_p(d.data_ptr()), 70, 45, C.c_float(1.0), None, C.c_float(9.5), C.c_uint64(11),
                                  C.c_void_p(out.data_ptr()), 3 * 70, None))
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), O.to_bitmap(img, 1.0, 11)[0])


@pytest.mark.gpu
def test_strip_owned_max_and_tonemap_assemble_the_whole_bitmap():
    """Image.toColors over several GPUs (Image.fs:37-50): the per-rank maxima over the rank's own strips meet in a max, and the
    per-rank bitmaps -- each byte written by exactly one rank -- meet in a bytewise max (what DistributedRenderer.to_bitmap
    does with two all-reduce(MAX)); the result equals the whole-frame bitmap and the oracle's (gamma 1: exactly)."""
    import ctypes as C

    import torch

    from fraytracer_b200 import _abi as A

    lib = ft.load_library()
    rng = np.random.default_rng(8)
    X, Y, sw = 83, 47, 8  # ragged: the last strip is 3 columns wide
    img = rng.uniform(0, 3, size=(X, Y, 3)).astype(np.float32)
    img[61, 5, 2] = 7.25  # the maximum lives in a strip of rank 61 // 8 % world
    d = torch.from_numpy(img).cuda()
    for world in (1, 2, 3):
        maxima, bitmaps = [], []
        for rank in range(world):
            mx = torch.zeros(1, dtype=torch.float32, device="cuda")
            A.check(lib.ft_image_max_strips_device(C.c_void_p(d.data_ptr()), X, Y, sw, rank, world, C.c_void_p(mx.data_ptr()), None))
            maxima.append(mx)
        owner = (61 // sw) % world
        assert float(maxima[owner].item()) == 7.25 and all(float(m.item()) < 7.25 for r, m in enumerate(maxima) if r != owner)
        gmax = torch.stack(maxima).max(dim=0).values  # all-reduce(MAX)
        for rank in range(world):
            out = torch.zeros((Y, X, 3), dtype=torch.uint8, device="cuda")
            A.check(lib.ft_tonemap_strips_device(C.c_void_p(d.data_ptr()), X, Y, sw, rank, world, C.c_float(1.0),
                                                 C.c_void_p(gmax.data_ptr()), C.c_float(0.0), C.c_uint64(11),
                                                 C.c_void_p(out.data_ptr()), 3 * X, None))
            bitmaps.append(out)
        torch.cuda.synchronize()
        # every byte position is written by at most one rank
        written = torch.stack([(b != 0) for b in bitmaps]).sum(dim=0)
        assert int(written.max().item()) <= 1
        whole = torch.stack(bitmaps).max(dim=0).values.cpu().numpy()  # all-reduce(MAX) over the bitmaps
        assert np.array_equal(whole, O.to_bitmap(img, 1.0, 11)[0])


@pytest.mark.gpu
def test_distributed_renderer_to_bitmap_single_rank():
    import torch

    from fraytracer_b200.distributed import DistributedRenderer

    w = scenes.default_scene(n_tori=200, size=120)
    tracer = ft.SceneTracer(w.scene, device=0)
    r = DistributedRenderer(tracer, w.size)
    frame = r.render(w.camera, w.epsilon, w.length)
    bitmap, mx = r.to_bitmap(frame, 1.0, 3)
    torch.cuda.synchronize()
    img = frame.cpu().numpy()
    want, want_max = O.to_bitmap(img, 1.0, 3)
    assert np.array_equal(bitmap.cpu().numpy(), want) and float(mx.item()) == want_max
