"""GPU parity tests proper: the CUDA path through the C ABI against the CPU oracle on the same inputs.
Bit-exact for colours (float32), hit object index, march step counts and shadow masks."""
import numpy as np
import pytest

import fraytracer_b200 as ft
from fraytracer_b200 import _abi as A
from fraytracer_b200 import scenes
from oracle import oracle as O

from conftest import bits

pytestmark = pytest.mark.gpu


def _check_render(w, X, Y):
    tracer = ft.SceneTracer(w.scene, device=0)
    size = ft.ImageSize(X, Y)
    rgb, leaf, stats = tracer.render(w.camera, size, w.epsilon, w.length, want_leaf=True)
    ref = O.OracleScene(tracer.description).render(w.camera, X, Y, w.epsilon, w.length, counters=True)
    assert stats["capped_rays"] == 0
    assert np.array_equal(leaf, ref["leaf"])
    assert np.array_equal(bits(rgb), bits(ref["rgb"]))
    assert stats["primary_rays"] == X * Y
    assert stats["primary_hits"] == ref["counters"]["primary_hits"]
    assert stats["shadow_rays"] == ref["counters"]["shadow_rays"]
    tracer.close()


def test_config1_default_scene_render_bit_exact():
    _check_render(scenes.default_scene(), 192, 192)


def test_config2_spheres_capsules_render_bit_exact():
    _check_render(scenes.spheres_capsules_scene(), 256, 144)


def test_config5_mixed_render_bit_exact():
    _check_render(scenes.mixed_scene(), 256, 144)


def test_non_square_and_ragged_sizes():
    w = scenes.default_scene(n_tori=120)
    for X, Y in ((1, 1), (7, 3), (33, 65), (130, 61)):
        _check_render(w, X, Y)


def test_trace_rays_matches_oracle_per_ray():
    w = scenes.default_scene()
    tracer = ft.SceneTracer(w.scene, device=0)
    orc = O.OracleScene(tracer.description)
    rng = np.random.default_rng(7)
    n = 20000
    origin = rng.normal(size=(n, 3)).astype(np.float32) * 6.0
    target = rng.normal(size=(n, 3)).astype(np.float32) * 2.0
    d = target - origin
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.concatenate([origin, d.astype(np.float32), np.full((n, 1), 25.0, np.float32),
                           rng.choice([0.01, 0.003, 0.05], size=(n, 1)).astype(np.float32)], axis=1)
    got, stats = tracer.trace_rays(rays)
    ref = orc.trace_rays(rays)
    assert np.array_equal(got["leaf"], ref["leaf"])
    assert np.array_equal(got["steps"], ref["steps"])
    assert np.array_equal(got["shadow"], ref["shadow"])
    assert np.array_equal(bits(got["rgb"]), bits(ref["rgb"]))
    assert (got["leaf"] >= 0).sum() > n // 10


def test_distance_matches_oracle():
    for w in (scenes.default_scene(), scenes.mixed_scene()):
        tracer = ft.SceneTracer(w.scene, device=0)
        orc = O.OracleScene(tracer.description)
        rng = np.random.default_rng(3)
        pts = (rng.normal(size=(50000, 3)) * 3.0).astype(np.float32)
        assert np.array_equal(bits(tracer.distance(pts)), bits(orc.distance(pts)))


def test_device_grid_equals_oracle_grid():
    w = scenes.mixed_scene()
    tracer = ft.SceneTracer(w.scene, device=0)
    g1, g2 = tracer.grids(), O.OracleScene(tracer.description).grids()
    assert len(g1) == len(g2) == 1
    for k in ("centers", "lower_bound"):
        assert np.array_equal(bits(g1[0][k]), bits(g2[0][k]))
    assert np.array_equal(g1[0]["item"], g2[0]["item"]) and np.array_equal(g1[0]["begin"], g2[0]["begin"])


def test_strip_partition_gather_equals_single_render():
    """multi-GPU path on one GPU: every 'rank' renders its strips into a compact buffer, the buffers are
    stacked as the all-gather would, the de-interleave kernel restores FColor[X,Y]; must be bit-identical."""
    import ctypes as C

    import torch

    from fraytracer_b200 import _abi as A
    from fraytracer_b200.distributed import strip_layout

    w = scenes.default_scene(n_tori=300)
    tracer = ft.SceneTracer(w.scene, device=0)
    for X, Y, sw, world in ((200, 120, 8, 2), (203, 77, 8, 4), (64, 64, 16, 3)):
        size = ft.ImageSize(X, Y)
        full, _, _ = tracer.render(w.camera, size, w.epsilon, w.length)
        spr, floats = strip_layout(X, Y, sw, world)
        gathered = torch.full((world, spr * sw, Y, 3), float("nan"), dtype=torch.float32, device="cuda:0")
        for r in range(world):
            tracer.render_device(w.camera, size, w.epsilon, w.length, gathered[r].data_ptr(), strip_width=sw,
                                 rank=r, world=world, stream=torch.cuda.current_stream().cuda_stream)
        final = torch.empty((X, Y, 3), dtype=torch.float32, device="cuda:0")
        A.check(tracer.lib.ft_deinterleave_device(C.c_void_p(gathered.data_ptr()), X, Y, sw, world, 3,
                                                  C.c_void_p(final.data_ptr()), None))
        torch.cuda.synchronize()
        assert not torch.isnan(gathered).any()  # padding columns are written too
        assert np.array_equal(bits(final.cpu().numpy()), bits(full))
        # the same in two chunks of local strips (what DistributedRenderer overlaps with the all-gather)
        final2 = torch.full((X, Y, 3), float("nan"), dtype=torch.float32, device="cuda:0")
        edges = [0, spr // 2, spr]
        for first, n in ((edges[0], edges[1] - edges[0]), (edges[1], edges[2] - edges[1])):
            if n == 0:
                continue
            chunk = torch.full((world, n * sw, Y, 3), float("nan"), dtype=torch.float32, device="cuda:0")
            for r in range(world):
                tracer.render_device(w.camera, size, w.epsilon, w.length, chunk[r].data_ptr(), strip_width=sw, rank=r,
                                     world=world, first_local_strip=first, n_local_strips=n,
                                     stream=torch.cuda.current_stream().cuda_stream)
            A.check(tracer.lib.ft_deinterleave_strips_device(C.c_void_p(chunk.data_ptr()), X, Y, sw, world, 3, first, n,
                                                             C.c_void_p(final2.data_ptr()), None))
        torch.cuda.synchronize()
        assert np.array_equal(bits(final2.cpu().numpy()), bits(full))


def test_fast_math_mode_within_tolerance():
    """FT_FLAG_FAST_MATH (FMA contraction) is not bit-exact.  The reference's normal is a forward difference
    with h = eps/8 (SdfForm.fs:106-115), which amplifies any change of rounding by ~1/h, so the achievable bar is:
    hit object index equal except < 1e-4 of the pixels (silhouettes); on equal-leaf hit pixels the relative colour
    error has median < 1e-4, 99.9th percentile < 5e-3, and < 1e-3 of them exceed 1e-2 (shadow-edge flips)."""
    from fraytracer_b200 import FT_FLAG_FAST_MATH

    for w, X, Y in ((scenes.default_scene(), 400, 400), (scenes.spheres_capsules_scene(), 480, 270)):
        tracer = ft.SceneTracer(w.scene, device=0)
        size = ft.ImageSize(X, Y)
        a, la, _ = tracer.render(w.camera, size, w.epsilon, w.length, want_leaf=True)
        b, lb, _ = tracer.render(w.camera, size, w.epsilon, w.length, want_leaf=True, flags=FT_FLAG_FAST_MATH)
        same = la == lb
        assert 1.0 - same.mean() < 1e-4
        sel = same & (la >= 0)
        rel = (np.abs(a - b) / np.maximum(np.abs(a), 1e-3)).max(axis=-1)[sel]
        assert np.median(rel) < 1e-4 and np.percentile(rel, 99.9) < 5e-3 and (rel > 1e-2).mean() < 1e-3
        assert np.array_equal(bits(a[la < 0]), bits(b[la < 0]))  # background pixels are exact


def test_escape_proof_changes_nothing_but_step_counts(monkeypatch):
    """rays resolved by the escape proof (render path, trace_rays without step output) must give the same
    colours, leaves and shadow masks as marching them to the end"""
    import ctypes as C

    from fraytracer_b200 import _abi as A

    for w in (scenes.default_scene(), scenes.mixed_scene(width=320, height=180)):
        size = ft.ImageSize(320, 180)
        fast = ft.SceneTracer(w.scene, device=0)
        monkeypatch.setenv("FT_NO_ESCAPE", "1")
        slow = ft.SceneTracer(w.scene, device=0)
        monkeypatch.delenv("FT_NO_ESCAPE")
        a, la, sa = fast.render(w.camera, size, w.epsilon, w.length, want_leaf=True)
        b, lb, sb = slow.render(w.camera, size, w.epsilon, w.length, want_leaf=True)
        assert sa["escaped_rays"] > 0 and sb["escaped_rays"] == 0
        assert np.array_equal(bits(a), bits(b)) and np.array_equal(la, lb)
        assert sa["shadow_rays"] == sb["shadow_rays"] and sa["primary_hits"] == sb["primary_hits"]
        # trace_rays without the step output uses the proof too
        rng = np.random.default_rng(11)
        n = 5000
        origin = (rng.normal(size=(n, 3)) * 8.0).astype(np.float32)
        d = rng.normal(size=(n, 3)).astype(np.float32)
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        rays = np.ascontiguousarray(np.concatenate([origin, d, np.full((n, 1), 40.0, np.float32),
                                                    np.full((n, 1), 0.01, np.float32)], axis=1), dtype=np.float32)
        full, _ = fast.trace_rays(rays)  # steps requested: exact marching
        rgb = np.empty((n, 3), np.float32)
        leaf = np.empty(n, np.int32)
        shadow = np.empty(n, np.uint32)
        st = A.ft_render_stats()
        A.check(fast.lib.ft_trace_rays(fast.handle, n, rays.ctypes.data_as(C.c_void_p), 0,
                                       rgb.ctypes.data_as(C.c_void_p), leaf.ctypes.data_as(C.c_void_p), None,
                                       shadow.ctypes.data_as(C.c_void_p), C.byref(st)))
        assert st.escaped_rays > 0
        assert np.array_equal(bits(rgb), bits(full["rgb"])) and np.array_equal(leaf, full["leaf"])
        assert np.array_equal(shadow, full["shadow"])


def test_watchdog_step_cap_reports_capped_rays():
    w = scenes.default_scene(n_tori=50)
    tracer = ft.SceneTracer(w.scene, device=0)
    tracer.set_step_cap(3)
    _, _, stats = tracer.render(w.camera, ft.ImageSize(64, 64), w.epsilon, w.length)
    assert stats["capped_rays"] > 0
    orc = O.OracleScene(tracer.description)
    orc.set_step_cap(3)
    ref = orc.render(w.camera, 64, 64, w.epsilon, w.length)
    rgb, _, _ = tracer.render(w.camera, ft.ImageSize(64, 64), w.epsilon, w.length)
    assert np.array_equal(bits(rgb), bits(ref["rgb"]))


def test_image_render_mirror_and_single_ray_closure():
    w = scenes.default_scene(n_tori=100)
    trace = ft.SdfScene.trace(w.scene)
    img = ft.Image.render(w.epsilon, w.length, ft.ImageSize(48, 32), w.camera, trace)
    assert img.shape == (48, 32, 3)
    orc = O.OracleScene(trace.description)
    ref = orc.render(w.camera, 48, 32, w.epsilon, w.length)
    assert np.array_equal(bits(img), bits(ref["rgb"]))
    ray = ft.Ray(Origin=(0, 0, -10), Direction=(0, 0, 1), Length=30.0, Epsilon=0.01)
    assert bits(np.array(trace(ray), np.float32)).tolist() == bits(orc.trace_rays(ray.as_row())["rgb"][0]).tolist()


def _nested_scene():
    from fraytracer_b200 import SdfForm, SdfObject, SdfScene

    P = SdfForm.Primitive
    b = scenes.SceneBuilder(seed=36)
    inner1 = SdfObject.union([b.randomSphere() for _ in range(7)])
    inner2 = SdfObject.union([b.randomCapsule() for _ in range(9)] + [b.randomTriangle() for _ in range(5)])
    outer = SdfObject.union([inner1, inner2, b.randomTorus(), b.randomTorus()])
    carved = SdfObject.subtract(SdfObject.intersect(outer, [P.sphere(Center=(0, 0, 0), Radius=3.0)]),
                                SdfForm.union([P.sphere(Center=(1, 1, 1), Radius=1.0), P.sphere(Center=(-1, 1, 0), Radius=0.7)]))
    return SdfScene(Object=carved, BackgroundColor=(0.05, 0.1, 0.2), Lights=scenes.default_lights())


def test_generic_interpreter_nested_composites_bit_exact():
    """nested unions, a union as the subtracted operand: not a chain program, runs the generic interpreter"""
    scene = _nested_scene()
    cam = scenes.default_camera()
    tracer = ft.SceneTracer(scene, device=0)
    assert tracer.info()["n_grids"] == 4
    rgb, leaf, _ = tracer.render(cam, ft.ImageSize(160, 120), 0.01, 30.0, want_leaf=True)
    ref = O.OracleScene(tracer.description).render(cam, 160, 120, 0.01, 30.0)
    assert np.array_equal(leaf, ref["leaf"]) and np.array_equal(bits(rgb), bits(ref["rgb"]))
    assert len(np.unique(leaf)) > 10


def test_union_smooth_within_tolerance():
    """SdfForm.unionSmooth uses MathF.Exp/Log (C runtime, not bit-reproducible): tolerance row.  Distances agree
    to 2e-6 absolute + 2e-5 relative with the oracle (glibc expf/logf vs CUDA expf/logf)."""
    from fraytracer_b200 import SdfForm, SdfMaterial, SdfObject, SdfScene

    P = SdfForm.Primitive
    forms = [P.sphere(Center=(0, 0, 0), Radius=1.0), P.sphere(Center=(1.5, 0, 0), Radius=0.8),
             P.torus(Center=(0, 1.5, 0), Normal=(0, 1, 0.2), MajorRadius=1.0, MinorRadius=0.25),
             P.capsule(From=(-1, -1, 0), To=(1, -1.5, 0.5), Radius=0.3)]
    obj = SdfObject.create(SdfMaterial.createSolid((0.8, 0.6, 0.4)), SdfForm.unionSmooth(0.25, forms))
    scene = SdfScene(Object=obj, BackgroundColor=(0.1, 0.1, 0.1), Lights=scenes.default_lights())
    tracer = ft.SceneTracer(scene, device=0)
    orc = O.OracleScene(tracer.description)
    pts = (np.random.default_rng(2).normal(size=(20000, 3)) * 2).astype(np.float32)
    assert np.allclose(tracer.distance(pts), orc.distance(pts), rtol=2e-5, atol=2e-6)
    cam = scenes.default_camera()
    rgb, leaf, _ = tracer.render(cam, ft.ImageSize(96, 96), 0.01, 30.0, want_leaf=True)
    ref = orc.render(cam, 96, 96, 0.01, 30.0)
    assert (leaf != ref["leaf"]).mean() < 0.01  # silhouette pixels only
    same = (leaf == ref["leaf"])
    rel = np.abs(rgb - ref["rgb"])[same] / np.maximum(np.abs(ref["rgb"])[same], 1e-3)
    assert np.percentile(rel, 99) < 2e-2


def test_chain_program_equals_generic_interpreter(monkeypatch):
    for w in (scenes.default_scene(), scenes.mixed_scene()):
        size = ft.ImageSize(240, 135)
        chain = ft.SceneTracer(w.scene, device=0)
        monkeypatch.setenv("FT_NO_CHAIN", "1")
        generic = ft.SceneTracer(w.scene, device=0)
        monkeypatch.delenv("FT_NO_CHAIN")
        a, la, _ = chain.render(w.camera, size, w.epsilon, w.length, want_leaf=True)
        b, lb, _ = generic.render(w.camera, size, w.epsilon, w.length, want_leaf=True)
        assert np.array_equal(bits(a), bits(b)) and np.array_equal(la, lb)


@pytest.mark.parametrize("make,stride", [(scenes.default_scene, 25), (scenes.spheres_capsules_scene, 48),
                                         (scenes.mixed_scene, 256)])
def test_full_size_frames_against_oracle_columns(make, stride):
    """BASELINE.json's full frame sizes (config 1: 1000x1000, config 2: 1920x1080, config 5: 7680x4320): every
    `stride`-th column of the GPU frame must equal the oracle bit for bit (the oracle finishes these columns in seconds)."""
    w = make()
    X, Y = w.size.X, w.size.Y
    tracer = ft.SceneTracer(w.scene, device=0)
    rgb, leaf, stats = tracer.render(w.camera, w.size, w.epsilon, w.length, want_leaf=True)
    ref = O.OracleScene(tracer.description).render(w.camera, X, Y, w.epsilon, w.length, x_stride=stride)
    cols = np.arange(0, X, stride)
    assert np.array_equal(bits(rgb[cols]), bits(ref["rgb"][cols]))
    assert np.array_equal(leaf[cols], ref["leaf"][cols])
    assert stats["capped_rays"] == 0 and stats["primary_rays"] == X * Y


def test_full_size_properties_config5_4k():
    """size-independent properties at a size the oracle cannot cover: (a) the frame is independent of how it is cut
    (one launch, two host bands, 8-column strips of 4 'ranks'); (b) re-rendering is idempotent; (c) the stats add up."""
    import ctypes as C

    import torch

    from fraytracer_b200 import _abi as A
    from fraytracer_b200.distributed import strip_layout

    w = scenes.mixed_scene(width=3840, height=2160)
    X, Y = w.size.X, w.size.Y
    tracer = ft.SceneTracer(w.scene, device=0)
    a, la, sa = tracer.render(w.camera, w.size, w.epsilon, w.length, want_leaf=True)
    b, lb, sb = tracer.render(w.camera, w.size, w.epsilon, w.length, want_leaf=True)
    assert np.array_equal(bits(a), bits(b)) and np.array_equal(la, lb)
    assert sa["primary_hits"] == int((la >= 0).sum()) == sb["primary_hits"]
    assert (bits(a[la < 0]) == bits(np.float32(0.1)).item()).all()  # misses carry the background colour exactly
    world, sw = 4, 8
    spr, _ = strip_layout(X, Y, sw, world)
    gathered = torch.empty((world, spr * sw, Y, 3), dtype=torch.float32, device="cuda:0")
    hits = 0
    for r in range(world):
        st = tracer.render_device(w.camera, w.size, w.epsilon, w.length, gathered[r].data_ptr(), strip_width=sw,
                                  rank=r, world=world, stream=torch.cuda.current_stream().cuda_stream, want_stats=True)
        hits += st["primary_hits"]
    final = torch.empty((X, Y, 3), dtype=torch.float32, device="cuda:0")
    A.check(tracer.lib.ft_deinterleave_device(C.c_void_p(gathered.data_ptr()), X, Y, sw, world, 3,
                                              C.c_void_p(final.data_ptr()), None))
    torch.cuda.synchronize()
    assert np.array_equal(bits(final.cpu().numpy()), bits(a)) and hits == sa["primary_hits"]


def test_fused_render_gather_equals_render():
    """ft_render_gather_device with world = 1 and several destinations: every destination holds ft_render's frame"""
    import torch

    w = scenes.default_scene(n_tori=200)
    tracer = ft.SceneTracer(w.scene, device=0)
    for X, Y in ((130, 61), (64, 64)):
        size = ft.ImageSize(X, Y)
        ref, _, _ = tracer.render(w.camera, size, w.epsilon, w.length)
        for world in (1, 3):
            dsts = [torch.full((X, Y, 3), -7.0, dtype=torch.float32, device="cuda") for _ in range(3)]
            for rank in range(world):
                tracer.render_gather_device(w.camera, size, w.epsilon, w.length, [d.data_ptr() for d in dsts],
                                            stream=torch.cuda.current_stream().cuda_stream, rank=rank, world=world)
            torch.cuda.synchronize()
            for d in dsts:
                assert np.array_equal(bits(d.cpu().numpy()), bits(ref))
    tracer.close()


def test_two_rank_p2p_gather_when_two_gpus_are_present():
    """the CUDA IPC path needs one process per GPU: run tools/check_p2p.py under torchrun when the box has >= 2 GPUs"""
    import os
    import subprocess
    import sys

    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (covered by `gpurun --gpus 2 -- torchrun ... tools/check_p2p.py`)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29517", os.path.join(root, "tools", "check_p2p.py")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_instrumented_kernel_counters_against_oracle_counters():
    """the roofline numerator comes from the instrumented (COUNT) kernel variant.  With per-ray step counts
    requested the escape proof is off, so its march-step total must equal the oracle's exactly, and its algorithmic flop
    total must lie between the oracle-derived lower bound (primitive evaluations + cell lookups + moves: the early exit
    and the inlined spheres only remove pruning tests) and the oracle-derived total of a scan without early exit."""
    for w in (scenes.default_scene(n_tori=300), scenes.spheres_capsules_scene(n_spheres=150, n_capsules=150)):
        tracer = ft.SceneTracer(w.scene, device=0)
        orc = O.OracleScene(tracer.description)
        X, Y = 160, 120
        cam = w.camera.to_c()
        # the frame's primary rays as explicit Ray records (Camera.uniformPixelToRay is part of the oracle's render)
        ref = orc.render(w.camera, X, Y, w.epsilon, w.length, counters=True)
        rgb, leaf, st = tracer.render(w.camera, ft.ImageSize(X, Y), w.epsilon, w.length, want_leaf=True,
                                      flags=A.FT_FLAG_COUNTERS)
        c = ref["counters"]
        assert st["primary_hits"] == c["primary_hits"] and st["shadow_rays"] == c["shadow_rays"]
        assert 0 < st["march_steps"] <= c["march_steps"]  # the escape proof and the shared first step only remove steps
        prims = 21 * c["eval_capsule"] + 27 * c["eval_torus"] + 85 * c["eval_triangle_edge"] + 34 * c["eval_triangle_face"]
        assert st["eval_flops"] > 0
        assert st["eval_flops"] <= prims + 10 * c["eval_sphere"] + 19 * c["union_lookups"] + 2 * c["lower_bound_tests"] + \
            13 * c["bound_tests"] + 7 * c["march_steps"] + 200 * c["primary_rays"]
        # per-ray step counts (no escape proof): exact
        rays = np.zeros((X * Y, 8), np.float32)
        rays[:, 0:3] = np.array(cam.position)
        xs, ys = np.meshgrid(np.arange(X), np.arange(Y), indexing="ij")
        m = np.float32(max(X, Y))
        u = (xs.astype(np.float32) / m - np.float32(0.5))[..., None]
        v = (ys.astype(np.float32) / m - np.float32(0.5))[..., None]
        d = (np.array(cam.forward, np.float32) + np.array(cam.right_scaled, np.float32) * u) + np.array(cam.up_scaled, np.float32) * v
        rays[:, 3:6] = (d / np.sqrt(((d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]))[..., None]).reshape(-1, 3)
        rays[:, 6], rays[:, 7] = w.length, w.epsilon
        got, st2 = tracer.trace_rays(rays, flags=A.FT_FLAG_COUNTERS)
        oref = orc.trace_rays(rays, counters=True)
        assert np.array_equal(got["steps"], oref["steps"]) and np.array_equal(bits(got["rgb"]), bits(oref["rgb"]))
        assert st2["march_steps"] == oref["counters"]["march_steps"]
        c2 = oref["counters"]
        prims2 = 21 * c2["eval_capsule"] + 27 * c2["eval_torus"] + 85 * c2["eval_triangle_edge"] + 34 * c2["eval_triangle_face"]
        assert st2["eval_flops"] >= prims2 + 7 * (c2["march_steps"] - c2["primary_hits"] - c2["shadow_rays"])
        tracer.close()


def test_extension_strip_rendering_equals_whole_frame():
    """EXTENSIONS: the strips of 3 'ranks' rendered one after the other into two destination framebuffers reproduce the
    whole-frame spectral / cloud render bit for bit (the multi-GPU path minus the second GPU)"""
    import torch

    from fraytracer_b200 import CloudParams

    cases = (("spectral", scenes.spectral_glass_scene(width=150, height=90), dict(n_wavelengths=8, max_bounces=5, seed=3)),
             ("cloud", scenes.cloud_scene(width=150, height=90), dict(samples_per_pixel=5, seed=3)))
    for kind, w, kw in cases:
        tracer = ft.SceneTracer(w.scene, device=0)
        if kind == "spectral":
            ref, st = tracer.render_spectral(w.camera, w.size, w.epsilon, w.length, 8, 5, 3)
        else:
            ref, st = tracer.render_cloud(w.camera, w.size, w.epsilon, w.length, CloudParams(), 5, 3)
        X, Y = w.size.X, w.size.Y
        for world, sw in ((1, 8), (3, 8), (2, 20), (3, 0), (1, 0)):
            dsts = [torch.full((X, Y, 3), -7.0, dtype=torch.float32, device="cuda") for _ in range(2)]
            rays = 0
            for rank in range(world):
                s = tracer.render_extension_gather_device(kind, w.camera, w.size, w.epsilon, w.length, [d.data_ptr() for d in dsts],
                                                          strip_width=sw, rank=rank, world=world, **kw)
                rays += s["primary_rays"] + s["shadow_rays"]
            torch.cuda.synchronize()
            for d in dsts:
                assert np.array_equal(bits(d.cpu().numpy()), bits(ref)), (kind, world, sw)
            assert rays == st["primary_rays"] + st["shadow_rays"]
        tracer.close()


@pytest.mark.gpu
@pytest.mark.parametrize("make", [lambda: scenes.default_scene(n_tori=300, size=160),
                                  lambda: scenes.spheres_capsules_scene(width=240, height=135),
                                  lambda: scenes.mixed_scene(width=256, height=144)])
def test_subgrid_certified_path_equals_sequential_scan(make, monkeypatch):
    """The nearest-candidate sub-grid (DESIGN.md "sub-grid") against the reference's sequential candidate scan
    (SdfForm.fs:27-33) on the same device: (a) sub-grid off, (b) sub-grid on, (c) sub-grid built but detached by
    ft_scene_set_debug bit 0 ("distrust every certificate": the host clears sub_valid and the fast region in the kernel
    parameters, so every evaluation of the same handle runs the scan) -- frames, leaf ids, per-ray step counts and
    Distance values must be identical bit for bit; (b) must also equal the oracle."""
    w = make()
    X, Y = w.size.X, w.size.Y
    monkeypatch.setenv("FT_NO_SUBGRID", "1")
    plain = ft.SceneTracer(w.scene, device=0)
    monkeypatch.delenv("FT_NO_SUBGRID")
    assert plain.subgrid_info()["resolution"] == 0
    sub = ft.SceneTracer(w.scene, device=0)
    info = sub.subgrid_info()
    assert info["resolution"] in (1, 2, 4, 8) and info["n_entries"] >= info["n_subcells"] > 0
    rng = np.random.default_rng(5)
    pts = (rng.random((20000, 3), dtype=np.float32) - 0.5) * 12.0  # inside and outside the lookup's box
    rays = np.zeros((4096, 8), np.float32)
    rays[:, 0:3] = (rng.random((4096, 3), dtype=np.float32) - 0.5) * 14.0
    d = -rays[:, 0:3] + (rng.random((4096, 3), dtype=np.float32) - 0.5) * 4.0
    rays[:, 3:6] = d / np.linalg.norm(d, axis=1, keepdims=True)
    rays[:, 6], rays[:, 7] = 30.0, 0.01
    ref = O.OracleScene(sub.description).render(w.camera, X, Y, w.epsilon, w.length)
    results = []
    for tracer, debug in ((plain, 0), (sub, 0), (sub, 1)):
        tracer.set_debug(debug)
        rgb, leaf, _ = tracer.render(w.camera, w.size, w.epsilon, w.length, want_leaf=True)
        tr, _ = tracer.trace_rays(rays)
        results.append((bits(rgb), leaf, tr["steps"], bits(tr["rgb"]), tr["leaf"], bits(tracer.distance(pts))))
    for r in results[1:]:
        for a, b in zip(results[0], r):
            assert np.array_equal(a, b)
    assert np.array_equal(results[1][0], bits(ref["rgb"])) and np.array_equal(results[1][1], ref["leaf"])
    assert np.array_equal(results[1][5], bits(O.OracleScene(sub.description).distance(pts)))


@pytest.mark.gpu
def test_tile_order_hints_change_nothing_but_the_schedule():
    """ft_api.cu TileHint / tile_rank.cu: from the third frame of a shape on, tiles are drawn in an order ranked from the costs
    an earlier frame recorded (most expensive first; a timed merge of expensive and cheap tiles for host destinations).  Pixels are
    independent of the order tiles are drawn in: every path (staged ft_render, zero-copy ft_render, device strips, fused gather)
    must return the first frame's bits again and again, and the frame rendered with hints disabled."""
    import ctypes as C

    import torch

    w = scenes.spheres_capsules_scene(width=333, height=251)  # ragged: partial tiles at both edges
    X, Y = w.size.X, w.size.Y
    tracer = ft.SceneTracer(w.scene, device=0)
    ref = O.OracleScene(tracer.description).render(w.camera, X, Y, w.epsilon, w.length)
    host = C.c_void_p()
    A.check(tracer.lib.ft_host_alloc(C.byref(host), X * Y * 12))
    mapped = np.ctypeslib.as_array(C.cast(host, C.POINTER(C.c_float)), shape=(X, Y, 3))
    dev = torch.empty((X, Y, 3), dtype=torch.float32, device="cuda:0")
    dst = [torch.zeros((X, Y, 3), dtype=torch.float32, device="cuda:0") for _ in range(2)]
    for frame in range(6):  # frames 0 and 1 run unhinted and measure, frames 2.. are ordered by the costs of frame N - 2
        tracer.set_debug(4 if frame == 5 else 0)  # the last frame: hints disabled again
        rgb, leaf, _ = tracer.render(w.camera, w.size, w.epsilon, w.length, want_leaf=True)  # pageable: two staged bands
        assert np.array_equal(bits(rgb), bits(ref["rgb"])) and np.array_equal(leaf, ref["leaf"]), frame
        mapped.fill(-1.0)
        tracer.render(w.camera, w.size, w.epsilon, w.length, out=mapped)  # zero-copy, whole-tile stores
        assert np.array_equal(bits(mapped), bits(ref["rgb"])), frame
        dev.fill_(-1.0)
        tracer.render_device(w.camera, w.size, w.epsilon, w.length, dev.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
        for d in dst:
            d.fill_(-1.0)
        tracer.render_gather_device(w.camera, w.size, w.epsilon, w.length, [d.data_ptr() for d in dst],
                                    stream=torch.cuda.current_stream().cuda_stream, strip_width=8)
        torch.cuda.synchronize()
        assert np.array_equal(bits(dev.cpu().numpy()), bits(ref["rgb"])), frame
        for d in dst:
            assert np.array_equal(bits(d.cpu().numpy()), bits(ref["rgb"])), frame
    A.check(tracer.lib.ft_host_free(host))


@pytest.mark.gpu
@pytest.mark.parametrize("size", [(64, 72), (128, 32), (8, 8), (4, 8), (400, 248), (62, 72), (64, 70)])
def test_zero_copy_frames_equal_the_oracle(size):
    """ft_render into a mapped host frame (RenderTileJob: whole tiles leave as 128-bit stores): frames made of complete 4 x 8 tiles
    (column heights of 9, 4, 1 and 31 tiles; a single tile) and frames with partial tiles at either edge.  Every frame of a
    sequence (natural order, then ranked orders) must equal the oracle bit for bit, with no pixel left unwritten."""
    import ctypes as C

    X, Y = size
    w = scenes.spheres_capsules_scene(width=X, height=Y)
    tracer = ft.SceneTracer(w.scene, device=0)
    ref = O.OracleScene(tracer.description).render(w.camera, X, Y, w.epsilon, w.length)
    host = C.c_void_p()
    A.check(tracer.lib.ft_host_alloc(C.byref(host), X * Y * 12))
    mapped = np.ctypeslib.as_array(C.cast(host, C.POINTER(C.c_float)), shape=(X, Y, 3))
    for frame in range(5):
        mapped.fill(np.nan)
        tracer.render(w.camera, w.size, w.epsilon, w.length, out=mapped)
        assert np.array_equal(bits(mapped), bits(ref["rgb"])), frame
    A.check(tracer.lib.ft_host_free(host))
    tracer.close()


@pytest.mark.gpu
@pytest.mark.parametrize("make", [scenes.default_scene, scenes.spheres_capsules_scene])
def test_hinted_frames_equal_the_first_frame_at_full_size(make):
    """BASELINE frame sizes (1000x1000, 1920x1080): every later frame of a shape -- drawn in ranked tile order, a permutation that
    tile_rank.cu derives on the device -- equals the first frame (natural order) bit for bit, leaf indices included, for a device
    destination (longest-processing-time-first) and for a mapped host destination (timed merge)."""
    import ctypes as C

    import torch

    w = make()
    X, Y = w.size.X, w.size.Y
    tracer = ft.SceneTracer(w.scene, device=0)
    host = C.c_void_p()
    A.check(tracer.lib.ft_host_alloc(C.byref(host), X * Y * 12))
    mapped = np.ctypeslib.as_array(C.cast(host, C.POINTER(C.c_float)), shape=(X, Y, 3))
    dev = torch.empty((X, Y, 3), dtype=torch.float32, device="cuda:0")
    first_dev = first_host = None
    for frame in range(5):
        dev.fill_(-1.0)
        tracer.render_device(w.camera, w.size, w.epsilon, w.length, dev.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        mapped.fill(-1.0)
        tracer.render(w.camera, w.size, w.epsilon, w.length, out=mapped)
        if frame == 0:
            first_dev, first_host = dev.cpu().numpy().copy(), mapped.copy()
            assert np.array_equal(bits(first_dev), bits(first_host))
        else:
            assert np.array_equal(bits(dev.cpu().numpy()), bits(first_dev)), frame
            assert np.array_equal(bits(mapped), bits(first_host)), frame
    A.check(tracer.lib.ft_host_free(host))
    tracer.close()
