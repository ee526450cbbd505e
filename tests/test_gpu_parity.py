"""GPU parity tests: the sm_100a path, called through the C ABI, against the CPU oracle.

Bars (SURVEY §8c / north_star):
  * closest hit of caller-given rays: object id EQUAL and t within 1e-5 relative (the FP64 refine
    makes it bit-identical in practice; that stronger property is asserted too);
  * scatter / emitted with caller-given random inputs: FP32 shading within 2e-4 of the double oracle;
  * converged images: RMSE(gpu, oracle) <= 1.25 * sqrt((R_gg^2 + R_oo^2) / 2) where R_xx is the
    RMSE between two renders of the same implementation with different seeds (Monte-Carlo noise);
  * gamma / 8-bit quantise (core.clj:52-57): bit-exact given the same float sums.
"""
import itertools
import math
import random

import numpy as np
import pytest

import oracle
import raytrace_clj_b200 as rt
from raytrace_clj_b200.util import vec3

from helpers import camera_block0, camera_rays, u01

pytestmark = pytest.mark.gpu

FMAX = float(np.finfo(np.float32).max)
REL_TOL_T = 1e-5   # north_star: primary-ray t within 1e-5 relative


@pytest.fixture(scope="module")
def renderer():
    r = rt.native.Renderer([0])
    yield r
    r.close()


@pytest.fixture(scope="module")
def scene_c2(random_scene_flat):
    flat, cam_type, cam = random_scene_flat
    return flat, cam_type, cam, oracle.Scene(flat)


def _assert_hits_match(t_gpu, id_gpu, t_ref, id_ref):
    assert np.array_equal(id_gpu, id_ref), f"{int((id_gpu != id_ref).sum())} object ids differ"
    hit = id_ref >= 0
    assert np.all(np.isinf(t_gpu[~hit]))
    rel = np.abs(t_gpu[hit] - t_ref[hit]) / np.abs(t_ref[hit])
    assert rel.max(initial=0.0) <= REL_TOL_T
    # stronger: the FP64 refine follows the reference's operation order, so t is bit-identical
    assert np.array_equal(t_gpu[hit], t_ref[hit])


def test_trace_primary_camera_rays_c2(renderer, scene_c2):
    flat, cam_type, cam, S = scene_c2
    renderer.set_scene(flat)
    o, d, tm = camera_rays(cam, 1200, 800, 200_000, np.random.default_rng(11))
    t_gpu, id_gpu = renderer.trace_primary(o, d, tm, 0.001, FMAX)
    t_ref, id_ref = S.hit(o, d, tm, 0.001, FMAX)
    _assert_hits_match(t_gpu, id_gpu, t_ref, id_ref)
    assert (id_ref >= 0).all()          # every primary ray ends on something (the sky dome encloses the scene)


def test_trace_secondary_like_rays(renderer, scene_c2):
    """Rays that start ON surfaces (t_min = 0.001 self-hit rejection), inside spheres, and far outside."""
    flat, cam_type, cam, S = scene_c2
    renderer.set_scene(flat)
    rng = np.random.default_rng(5)
    o, d, tm = camera_rays(cam, 1200, 800, 60_000, rng)
    t, ids = S.hit(o, d, tm, 0.001, FMAX)
    p = (o.astype(np.float64) + t[:, None] * d.astype(np.float64)).astype(np.float32)   # hit points as float32
    nd = rng.normal(size=p.shape).astype(np.float32)
    nd *= rng.uniform(0.05, 2.0, size=(len(p), 1)).astype(np.float32)                   # |d| in (0, 2) like Lambert
    t_gpu, id_gpu = renderer.trace_primary(p, nd, tm, 0.001, FMAX)
    t_ref, id_ref = S.hit(p, nd, tm, 0.001, FMAX)
    _assert_hits_match(t_gpu, id_gpu, t_ref, id_ref)
    # origins anywhere in the scene volume (many start inside the ground sphere or small spheres)
    o2 = rng.uniform(-15, 15, size=(60_000, 3)).astype(np.float32)
    o2[:, 1] = rng.uniform(-1, 3, size=60_000)
    d2 = rng.normal(size=o2.shape).astype(np.float32)
    tm2 = rng.random(60_000).astype(np.float32)
    t_gpu, id_gpu = renderer.trace_primary(o2, d2, tm2, 0.001, FMAX)
    t_ref, id_ref = S.hit(o2, d2, tm2, 0.001, FMAX)
    _assert_hits_match(t_gpu, id_gpu, t_ref, id_ref)


def test_trace_grazing_rays(renderer, scene_c2):
    """Rays aimed at sphere silhouettes (discriminant ~ 0): the FP32 cull must not lose any hit."""
    flat, cam_type, cam, S = scene_c2
    renderer.set_scene(flat)
    rng = np.random.default_rng(9)
    n = 80_000
    k = rng.integers(0, flat.n_spheres, n)
    c = flat.center0_r[k, :3].astype(np.float64)
    r = flat.center0_r[k, 3].astype(np.float64)
    o = np.tile(np.array([13.0, 2.0, 3.0]), (n, 1)) + rng.normal(scale=0.5, size=(n, 3))
    to_c = c - o
    dist = np.linalg.norm(to_c, axis=1)
    perp = np.cross(to_c, rng.normal(size=(n, 3)))
    perp /= np.linalg.norm(perp, axis=1)[:, None]
    # aim at the silhouette +- a few 1e-7 relative
    target = c + perp * (r * (1.0 + rng.normal(scale=3e-7, size=n)))[:, None]
    d = (target - o) * rng.uniform(0.2, 3.0, size=(n, 1)) / dist[:, None]
    tm = np.zeros(n, np.float32)
    o32, d32 = o.astype(np.float32), d.astype(np.float32)
    t_gpu, id_gpu = renderer.trace_primary(o32, d32, tm, 0.001, FMAX)
    t_ref, id_ref = S.hit(o32, d32, tm, 0.001, FMAX)
    _assert_hits_match(t_gpu, id_gpu, t_ref, id_ref)


def _grazing_rays(flat, rng, n, origin_scale, jitter):
    """n rays from origins |o| ~ origin_scale aimed at sphere silhouettes +- jitter (relative)."""
    k = rng.integers(0, flat.n_spheres, n)
    c = flat.center0_r[k, :3].astype(np.float64)
    r = flat.center0_r[k, 3].astype(np.float64)
    o = rng.normal(size=(n, 3)) * origin_scale
    to_c = c - o
    dist = np.linalg.norm(to_c, axis=1)
    perp = np.cross(to_c, rng.normal(size=(n, 3)))
    perp /= np.linalg.norm(perp, axis=1)[:, None]
    target = c + perp * (r * (1.0 + rng.normal(scale=jitter, size=n)))[:, None]
    d = (target - o) * rng.uniform(0.2, 3.0, size=(n, 1)) / dist[:, None]
    return o.astype(np.float32), d.astype(np.float32)


def test_cull_never_under_reports(renderer, scene_c2):
    """The FP32 cull (expanded-form key, Culler in rt_kernels.cuh) may over-report but must never lose a
    (ray, sphere) pair the exact FP64 test accepts — checked pair by pair on the device over ~3e8 pairs:
    camera rays, rays from surfaces, silhouette-grazing rays, and origins far from the scene (where the
    expanded form cancels worst: |o| up to ~3000)."""
    flat, cam_type, cam, S = scene_c2
    renderer.set_scene(flat)
    rng = np.random.default_rng(21)
    families = {}
    o, d, tm = camera_rays(cam, 1200, 800, 100_000, rng)
    families["camera"] = (o, d, tm)
    t, _ = S.hit(o, d, tm, 0.001, FMAX)
    p = (o.astype(np.float64) + t[:, None] * d.astype(np.float64)).astype(np.float32)
    nd = rng.normal(size=p.shape).astype(np.float32) * rng.uniform(0.05, 2.0, size=(len(p), 1)).astype(np.float32)
    families["surface"] = (p, nd, tm)
    for scale in (1.0, 10.0, 100.0, 1000.0):
        for jitter in (3e-7, 1e-4):
            og, dg = _grazing_rays(flat, rng, 50_000, scale, jitter)
            families[f"grazing |o|~{scale:g} +-{jitter:g}"] = (og, dg, rng.random(len(og)).astype(np.float32))
    o2 = rng.uniform(-15, 15, size=(100_000, 3)).astype(np.float32)
    o2[:, 1] = rng.uniform(-1, 3, size=len(o2))
    families["volume"] = (o2, rng.normal(size=o2.shape).astype(np.float32), rng.random(len(o2)).astype(np.float32))
    total_pairs = 0
    for name, (o, d, tm) in families.items():
        for tmin in (0.0, 0.001):
            lost, survivors, candidates = renderer.cull_check(o, d, tm, tmin, FMAX)
            assert lost == 0, f"{name}: the cull lost {lost} of {candidates} exact candidates (tmin {tmin})"
            assert survivors >= candidates
        total_pairs += len(o) * flat.n_spheres
        if name == "camera":   # selectivity: the cull passes few pairs beyond the true candidates
            assert survivors <= 1.5 * candidates + 0.002 * len(o) * flat.n_spheres
    assert total_pairs > 2e8
    with pytest.raises(rt.native.NativeError):
        renderer.cull_check(o, d, tm, -1.0, FMAX)


# hitable_test.clj:8-19
GRIDPOINTS = [25.0 * np.array(p, float) for p in itertools.product((-1, 0, 1), repeat=3)]
DIRECTIONS = [5.0 * np.array(p, float) for p in itertools.product((-1, 0, 1), repeat=3) if any(p)]


@pytest.mark.parametrize("moving", [False, True])
def test_reference_known_answers_on_gpu(renderer, moving):
    """hitable_test.clj:23-47 / 61-83 through the CUDA path: hit from c+dir along -dir, miss along +dir,
    grazing rays (discriminant exactly 0), ray from the centre; t_min = 0 like the reference test."""
    m = rt.shader.lambertian(albedo=rt.texture.constant(color=vec3(.8, .8, .8)))
    for origin in GRIDPOINTS:
        if moving:
            s = rt.hitable.moving_sphere(center0=origin, t0=0.1, center1=origin + np.array([10., 20., 30.]), t1=0.9,
                                         radius=1.0, material=m)
            time = 0.1
        else:
            s = rt.hitable.sphere(center=origin, radius=1.0, material=m)
            time = 0.0
        flat = rt.native.marshal_world(rt.hitable.hitlist(items=[s]))
        renderer.set_scene(flat)
        S = oracle.Scene(flat)
        os_, ds_ = [], []
        for dr in DIRECTIONS:
            os_ += [origin + dr, origin + dr]
            ds_ += [-dr, dr]
        graze = [([1, 1, 0], [-1, 0, 0]), ([1, 1, 0], [0, -1, 0]), ([1, 0, 1], [0, 0, -1])]
        for off, dr in graze:
            os_.append(origin + np.array(off, float))
            ds_.append(np.array(dr, float))
        os_.append(origin)
        ds_.append(np.array([1.0, 1.0, 1.0]))
        o = np.array(os_, np.float32)
        d = np.array(ds_, np.float32)
        tm = np.full(len(o), time, np.float32)
        t_gpu, id_gpu = renderer.trace_primary(o, d, tm, 0.0, FMAX)
        t_ref, id_ref = S.hit(o, d, tm, 0.0, FMAX)
        _assert_hits_match(t_gpu, id_gpu, t_ref, id_ref)
        nd = len(DIRECTIONS)
        assert np.all(id_gpu[0:2 * nd:2] == 0) and np.all(id_gpu[1:2 * nd:2] == -1)
        for q, dr in enumerate(DIRECTIONS):
            assert t_gpu[2 * q] == pytest.approx(1.0 - 1.0 / np.linalg.norm(dr), rel=1e-12)
        if not moving:   # the float32 lerp inputs 0.1 / 0.9 make the moving grazing case inexact
            assert np.all(id_gpu[2 * nd:2 * nd + 3] == 0) and np.all(t_gpu[2 * nd:2 * nd + 3] == 1.0)
        assert id_gpu[-1] == 0 and t_gpu[-1] == pytest.approx(1 / math.sqrt(3), rel=1e-7)


def test_trace_edge_cases(renderer, scene_c2):
    flat, cam_type, cam, S = scene_c2
    renderer.set_scene(flat)
    t, ids = renderer.trace_primary(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32), None)
    assert len(t) == 0 and len(ids) == 0
    # ragged: 1 ray, 255, 257, 1025 (block and slot boundaries)
    for n in (1, 255, 257, 1025):
        o, d, tm = camera_rays(cam, 1200, 800, n, np.random.default_rng(n))
        _assert_hits_match(*renderer.trace_primary(o, d, tm), *S.hit(o, d, tm))
    # t-range: a tight t_max turns hits into misses exactly like the oracle (strict inequalities)
    o, d, tm = camera_rays(cam, 1200, 800, 5000, np.random.default_rng(1))
    _assert_hits_match(*renderer.trace_primary(o, d, tm, 0.5, 1.0), *S.hit(o, d, tm, 0.5, 1.0))
    # first-wins exact ties (hitable.clj:17-26) and a ray that hits nothing
    m = rt.shader.lambertian(albedo=rt.texture.constant(color=vec3(.5, .5, .5)))
    world = rt.hitable.hitlist(items=[rt.hitable.sphere(center=vec3(0, 0, -10), radius=1, material=m),
                                      rt.hitable.sphere(center=vec3(0, 0, -5), radius=1, material=m),
                                      rt.hitable.sphere(center=vec3(0, 0, -5), radius=1, material=m),
                                      rt.hitable.sphere(center=vec3(0, 0, 5), radius=1, material=m)])
    renderer.set_scene(rt.native.marshal_world(world))
    t, ids = renderer.trace_primary([[0, 0, 0], [0, 0, 0], [0, 0, 0]], [[0, 0, -1], [0, 1, 0], [0, 0, -4]], [0, 0, 0], 0.001)
    assert list(ids) == [1, -1, 1] and t[0] == 4.0 and math.isinf(t[1]) and t[2] == 1.0


def test_trace_tiled_scene_6000_spheres(renderer):
    """A scene larger than one shared-memory tile (the sphere list is streamed through smem)."""
    sc = rt.scene.make_scale_sweep_scene(640, 360, 6000, random.Random(5))
    flat = rt.native.marshal_world(sc["world"])
    assert flat.n_spheres > 4096
    cam_type, cam = rt.native.marshal_camera(sc["camera"])
    renderer.set_scene(flat)
    S = oracle.Scene(flat)
    o, d, tm = camera_rays(cam, 640, 360, 20_000, np.random.default_rng(2))
    _assert_hits_match(*renderer.trace_primary(o, d, tm), *S.hit(o, d, tm))


def test_generate_rays_matches_get_ray(renderer, scene_c2):
    """camera.clj:35-48 + core.clj:49-50 on the device vs the oracle's get-ray fed the same uniforms;
    the uniforms themselves vs a numpy Philox4x32-10."""
    flat, cam_type, cam, S = scene_c2
    renderer.set_camera(cam_type, cam)
    nx, ny = 1200, 800
    rng = np.random.default_rng(3)
    n = 4096
    ij = np.stack([rng.integers(0, nx, n), rng.integers(0, ny, n)], axis=1).astype(np.int32)
    s = rng.integers(0, 1024, n).astype(np.int32)
    o, d, tm, rnd = renderer.generate_rays(nx, ny, ij, s, seed=0x1234_5678_9ABC)
    blk = camera_block0(0x1234_5678_9ABC, ij[:, 1].astype(np.int64) * nx + ij[:, 0], s)
    assert np.array_equal(rnd[:, 0], u01(blk[:, 0])) and np.array_equal(rnd[:, 1], u01(blk[:, 1]))
    assert np.array_equal(rnd[:, 4], u01(blk[:, 2]))
    assert rnd[:, [0, 1, 4]].min() >= 0.0 and rnd[:, [0, 1, 4]].max() < 1.0
    for q in range(0, n, 16):
        u = (float(ij[q, 0]) + float(rnd[q, 0])) / nx
        v = (float(ij[q, 1]) + float(rnd[q, 1])) / ny
        ro, rd, rtm = oracle.get_ray(cam_type, cam, u, v, disk=(float(rnd[q, 2]), float(rnd[q, 3])), time_u=float(rnd[q, 4]))
        assert np.allclose(o[q], ro, atol=1e-6) and np.allclose(d[q], rd, rtol=0, atol=5e-6) and abs(tm[q] - rtm) < 1e-6
    # a lens with a real aperture: disk samples inside the unit disk, origin displaced in the (u, v) plane
    cam2 = cam.copy()
    cam2[21] = 0.5
    renderer.set_camera(cam_type, cam2)
    o, d, tm, rnd = renderer.generate_rays(nx, ny, ij, s, seed=9)
    assert np.all(rnd[:, 2] ** 2 + rnd[:, 3] ** 2 < 1.0) and np.abs(rnd[:, 2:4]).max() > 0.5
    for q in range(0, n, 64):
        u = (float(ij[q, 0]) + float(rnd[q, 0])) / nx
        v = (float(ij[q, 1]) + float(rnd[q, 1])) / ny
        ro, rd, rtm = oracle.get_ray(cam_type, cam2, u, v, disk=(float(rnd[q, 2]), float(rnd[q, 3])), time_u=float(rnd[q, 4]))
        assert np.allclose(o[q], ro, atol=2e-6) and np.allclose(d[q], rd, atol=5e-6)
    # pinhole camera (camera.clj:8-16): time 0, origin fixed
    pin = oracle.pinhole_camera([13, 2, 3], [0, 0, 0], [0, 1, 0], 20, 1.5).astype(np.float32)
    renderer.set_camera(rt.native.RT_CAM_PINHOLE, pin)
    o, d, tm, rnd = renderer.generate_rays(nx, ny, ij, s, seed=9)
    assert np.all(tm == 0) and np.allclose(o, pin[0:3])
    renderer.set_camera(cam_type, cam)


def _shade_inputs(renderer, S, flat, cam, n, seed):
    rng = np.random.default_rng(seed)
    o, d, tm = camera_rays(cam, 1200, 800, n, rng)
    t, ids = S.hit(o, d, tm)
    # second-bounce rays too, so every material is entered from outside and (for glass) from inside
    p = (o.astype(np.float64) + (t * (1 + 1e-3))[:, None] * d.astype(np.float64)).astype(np.float32)
    o = np.concatenate([o, p[: n // 2]])
    d = np.concatenate([d, d[: n // 2]])
    tm = np.concatenate([tm, tm[: n // 2]])
    t, ids = S.hit(o, d, tm)
    ball = rng.uniform(-1, 1, size=(len(o), 3))
    ball *= (rng.random(len(o)) ** (1 / 3) / np.linalg.norm(ball, axis=1))[:, None]
    return o, d, tm, t, ids, ball.astype(np.float32), rng.random(len(o)).astype(np.float32)


def test_shade_batch_matches_oracle(renderer, scene_c2):
    """shader.clj:29-119 + texture.clj:14-50 + hitable.clj:128-139,193-201 with caller-given randoms."""
    flat, cam_type, cam, S = scene_c2
    renderer.set_scene(flat)
    o, d, tm, t, ids, ball, u = _shade_inputs(renderer, S, flat, cam, 120_000, 21)
    g = renderer.shade_batch(o, d, tm, ids, t, ball, u)
    ref = S.shade_batch(o, d, tm, ids, ball, u)
    types = np.where(ids >= 0, flat.mat_type[flat.material_id[np.maximum(ids, 0)]], -1)
    assert set(np.unique(types)) >= {0, 1, 2, 3}
    same = g["flags"] == ref["flags"]
    # threshold cases (metal grazing reflection, schlick coin exactly at the boundary) may flip in FP32
    assert (~same).mean() < 2e-4, (~same).mean()
    for mt in (0, 1, 2, 3):
        sel = same & (types == mt)
        assert sel.sum() > 100
        assert np.allclose(g["emitted"][sel], ref["emitted"][sel], atol=2e-4), mt
        cont = sel & (ref["flags"] == 1)
        if mt == 2:   # reflect/refract choice near the Schlick threshold or the TIR boundary: compare per branch
            close = np.all(np.abs(g["dir"][cont] - ref["dir"][cont]) < 2e-3, axis=1)
            assert close.mean() > 0.999
            cont_idx = np.nonzero(cont)[0][close]
        else:
            cont_idx = np.nonzero(cont)[0]
        scale = np.maximum(1.0, np.linalg.norm(ref["dir"][cont_idx], axis=1))[:, None]
        assert np.all(np.abs(g["dir"][cont_idx] - ref["dir"][cont_idx]) <= 3e-4 * scale), mt
        assert np.allclose(g["origin"][cont_idx], ref["origin"][cont_idx], rtol=1e-6, atol=2e-4), mt
        # a checkerboard sample within ~1e-6 of a sine zero crossing may pick the other colour in FP32
        att_ok = np.all(np.abs(g["atten"][cont_idx] - ref["atten"][cont_idx]) <= 1e-6, axis=1)
        assert len(att_ok) == 0 or att_ok.mean() > 0.999, (mt, att_ok.mean())
    # checkerboard ground: both colours are produced and agree with the oracle
    ground = same & (ids == int(np.nonzero((flat.center0_r[:, 3] == 1000) & (flat.sphere_flags == 0))[0][0]))
    cols = {tuple(round(float(x), 3) for x in c) for c in g["atten"][ground][:2000]}
    assert (0.2, 0.3, 0.1) in cols and (0.9, 0.9, 0.9) in cols


def _rmse(a, b):
    return float(np.sqrt(np.mean((np.asarray(a, np.float64) - np.asarray(b, np.float64)) ** 2)))


def _render_parity(renderer, flat, cam_type, cam, nx, ny, ns, depth=50, variant=0):
    S = oracle.Scene(flat)
    renderer.set_scene(flat)
    renderer.set_camera(cam_type, cam)
    g1, img1 = renderer.render(nx, ny, ns, depth, seed=101, variant=variant)
    g2, _ = renderer.render(nx, ny, ns, depth, seed=202, variant=variant)
    g1, g2 = g1.astype(np.float64), g2.astype(np.float64)      # reduce in double (float32 sums over many pixels drift)
    o1 = S.render_accumulate(cam_type, cam, nx, ny, 0, ns, depth, seed=303)[0] / ns
    o2 = S.render_accumulate(cam_type, cam, nx, ny, 0, ns, depth, seed=404)[0] / ns
    r_gg, r_oo = _rmse(g1, g2), _rmse(o1, o2)
    cross = [_rmse(g, o) for g in (g1, g2) for o in (o1, o2)]
    bound = 1.25 * math.sqrt((r_gg ** 2 + r_oo ** 2) / 2)
    assert max(cross) <= bound, (cross, r_gg, r_oo)
    assert 0.8 < r_gg / r_oo < 1.25, (r_gg, r_oo)              # same noise level => same estimator variance
    # per-channel mean difference within 3 sigma / sqrt(P) of the per-pixel noise
    sigma = np.sqrt(((o1 - o2) ** 2).mean(axis=(0, 1)) / 2 + ((g1 - g2) ** 2).mean(axis=(0, 1)) / 2)
    mean_diff = np.abs((g1 + g2).mean(axis=(0, 1)) / 2 - (o1 + o2).mean(axis=(0, 1)) / 2)
    assert np.all(mean_diff <= 3 * sigma / math.sqrt(nx * ny) + 2e-4), (mean_diff, sigma)
    # PSNR of the gamma-2 8-bit images, reported for the record
    img_o = oracle.resolve(o1 * ns, ns)
    mse = np.mean((img1.astype(np.float64) - img_o.astype(np.float64)) ** 2)
    psnr = 10 * math.log10(255.0 ** 2 / mse)
    print(f"[parity] {nx}x{ny}x{ns}spp variant {variant}: RMSE cross {max(cross):.4f} <= bound {bound:.4f} "
          f"(R_gg {r_gg:.4f}, R_oo {r_oo:.4f}), PSNR(8-bit) {psnr:.1f} dB")
    return g1


@pytest.mark.parametrize("variant", [0, 1])
def test_render_matches_oracle_random_scene(renderer, scene_c2, variant):
    """BASELINE config 1 stand-in: make-random-scene 200x100, 100 spp (the `lein run out.ppm 200 100 100` shape)."""
    flat, cam_type, cam, _ = scene_c2
    sc = rt.scene.make_random_scene(200, 100, 11, True, random.Random(1))
    cam_type, cam = rt.native.marshal_camera(sc["camera"])
    _render_parity(renderer, flat, cam_type, cam, 200, 100, 100, variant=variant)


@pytest.mark.parametrize("variant", [0, 1])
def test_render_matches_oracle_defocus_camera(renderer, scene_c2, variant):
    """A lens with a real aperture (camera.clj:35-48 disk sample): camera rays no longer share an origin, so the
    wavefront's common-origin cull form must stay off and fresh camera rays join the general queue region."""
    flat, _, _, _ = scene_c2
    sc = rt.scene.make_random_scene(160, 100, 11, True, random.Random(1))
    cam_type, cam = rt.native.marshal_camera(sc["camera"])
    cam = np.array(cam, np.float32)
    cam[21] = 0.3          # aperture (lens radius 0.15), focus distance 10 already baked into lleft / horiz / vert
    _render_parity(renderer, flat, cam_type, cam, 160, 100, 96, variant=variant)


@pytest.mark.parametrize("variant", [0, 1])
def test_render_matches_oracle_material_stress(renderer, variant):
    """BASELINE config 4: metal/glass-heavy mix, depth 50 (long paths, absorbed rays, TIR)."""
    sc = rt.scene.make_material_stress_scene(160, 96, 11, random.Random(4))
    flat = rt.native.marshal_world(sc["world"])
    cam_type, cam = rt.native.marshal_camera(sc["camera"])
    _render_parity(renderer, flat, cam_type, cam, 160, 96, 96, variant=variant)


@pytest.mark.parametrize("variant", [0, 1])
def test_render_matches_oracle_tiled_scene(renderer, variant):
    """BASELINE config 5's shape at test scale: more spheres than one shared-memory tile holds (> 4096), so the
    cull streams the sphere list through shared memory tile by tile and flushes the survivors chunk by chunk."""
    sc = rt.scene.make_scale_sweep_scene(96, 64, 5000, random.Random(5))
    flat = rt.native.marshal_world(sc["world"])
    assert flat.n_spheres > 4096
    cam_type, cam = rt.native.marshal_camera(sc["camera"])
    _render_parity(renderer, flat, cam_type, cam, 96, 64, 48, variant=variant)


@pytest.mark.parametrize("variant", [0, 1])
def test_render_matches_oracle_two_spheres_and_depth_cutoff(renderer, variant):
    """make-two-spheres (scene.clj:9-49: UVGradient on a Lambertian UVSphere) and a depth cutoff of 2."""
    sc = rt.scene.make_two_spheres(120, 80)
    flat = rt.native.marshal_world(sc["world"])
    cam_type, cam = rt.native.marshal_camera(sc["camera"])
    _render_parity(renderer, flat, cam_type, cam, 120, 80, 128, variant=variant)
    _render_parity(renderer, flat, cam_type, cam, 120, 80, 128, depth=2, variant=variant)
    # depth 0: only emission of the first hit (core.clj:26 (pos? depth) fails at once)
    lin, _ = renderer.render(120, 80, 16, 0, seed=1, variant=variant)
    S = oracle.Scene(flat)
    ref = S.render_accumulate(cam_type, cam, 120, 80, 0, 16, 0, seed=2)[0] / 16
    assert _rmse(lin, ref) < 0.05 and abs(lin.mean() - ref.mean()) < 5e-3


@pytest.mark.parametrize("variant", [0, 1])
def test_counters_define_the_metric(renderer, scene_c2, variant):
    """samples == nx*ny*ns; tests == rays * N (brute force, counted on device); every path ends once."""
    flat, cam_type, cam, S = scene_c2
    renderer.set_scene(flat)
    renderer.set_camera(cam_type, cam)
    renderer.reset_counters()
    nx, ny, ns = 150, 100, 8
    renderer.render(nx, ny, ns, 50, seed=5, variant=variant)
    c = renderer.counters()
    assert c["samples"] == nx * ny * ns
    assert c["sphere_tests"] == c["rays"] * flat.n_spheres
    assert c["term_light"] + c["term_absorb"] + c["term_depth"] + c["term_miss"] == c["samples"]
    assert c["kernel_ns"] > 0 and c["candidates"] >= c["rays"]          # the sky dome is always a candidate
    _, oc = S.render_accumulate(cam_type, cam, nx, ny, 0, ns, 50, seed=6)
    assert abs(c["rays"] / c["samples"] - oc["rays"] / oc["samples"]) < 0.05   # same mean path length
    assert abs(c["term_light"] - oc["term_light"]) < 0.01 * c["samples"]


@pytest.mark.parametrize("variant", [0, 1])
def test_render_is_reproducible(renderer, scene_c2, variant):
    """The same seed renders the same paths whatever the scheduling: every counter (rays, samples, how each path ended)
    is identical across runs and the float sums differ only by the order of the atomic adds.  (A race in the queue
    logic — a lost or duplicated path, a stale closest hit — would show here; compute-sanitizer is closed on this pool.)"""
    flat, cam_type, cam, _ = scene_c2
    renderer.set_scene(flat)
    sc = rt.scene.make_random_scene(480, 320, 11, True, random.Random(1))
    cam_type, cam = rt.native.marshal_camera(sc["camera"])
    renderer.set_camera(cam_type, cam)
    runs = []
    for _ in range(3):
        renderer.reset_counters()
        lin, img = renderer.render(480, 320, 4, 50, seed=99, variant=variant)     # 614 400 samples: two lanes + tail
        c = renderer.counters()
        runs.append((lin, img, {k: c[k] for k in ("rays", "samples", "term_light", "term_absorb", "term_depth", "term_miss", "candidates")}))
    for lin, img, c in runs[1:]:
        assert c == runs[0][2]
        assert np.allclose(lin, runs[0][0], rtol=1e-5, atol=1e-6)
        assert (img != runs[0][1]).mean() < 1e-4          # an 8-bit value may sit on a rounding edge
    c = runs[0][2]
    assert c["samples"] == 480 * 320 * 4 == c["term_light"] + c["term_absorb"] + c["term_depth"] + c["term_miss"]


def test_capped_megakernel_is_reproducible(scene_c2):
    """Round 1's register-capped megakernel rendered one path in ~600 k differently from run to run.  Cause: the shading
    code was inlined once per path slot, the copies rounded differently, and the slot a (pixel, sample) lands in depends
    on the dynamic work counter — not a race.  With ONE out-of-line shading body (mega_shade_one) the capped build is
    as reproducible as the others: identical counters over 12 runs, and equal to the uncapped build and the wavefront."""
    flat, _, _, _ = scene_c2
    sc = rt.scene.make_random_scene(480, 320, 11, True, random.Random(1))
    cam_type, cam = rt.native.marshal_camera(sc["camera"])
    keys = ("rays", "samples", "term_light", "term_absorb", "term_depth", "term_miss")
    seen = {}
    with rt.native.Renderer([0]) as r:
        r.set_scene(flat)
        r.set_camera(cam_type, cam)
        for cap, variant, runs in ((1, 0, 12), (0, 0, 3), (0, 1, 3)):
            r.set_option("mega_regcap", cap)
            for _ in range(runs):
                r.reset_counters()
                r.render(480, 320, 4, 50, seed=99, variant=variant, linear=False, rgb8=False)
                c = r.counters()
                seen.setdefault((cap, variant), set()).add(tuple(c[k] for k in keys))
    assert all(len(v) == 1 for v in seen.values()), seen             # every build reproduces itself run after run
    assert seen[(1, 0)] == seen[(0, 0)]                                # capped == uncapped: the same shading body
    mega, wave = next(iter(seen[(0, 0)])), next(iter(seen[(0, 1)]))    # the wavefront inlines its own copy: FP32 threshold flips only
    assert abs(mega[0] - wave[0]) <= 1e-5 * wave[0] and mega[1] == wave[1]


@pytest.mark.parametrize("variant", [0, 1])
def test_resolve_bit_exact_and_sharding(renderer, scene_c2, variant):
    """rt_render_accumulate_device + rt_resolve_device on caller-owned device buffers (the per-GPU leg of the
    sharded render): sample slices and interleaved rows add up to the unsharded sums; the 8-bit resolve is
    bit-identical to core.clj:52-57 evaluated by the oracle on the same float sums."""
    import torch

    flat, cam_type, cam, S = scene_c2
    renderer.set_scene(flat)
    renderer.set_camera(cam_type, cam)
    nx, ny, ns = 160, 90, 8
    dev = torch.device("cuda:0")
    full = torch.zeros(ny, nx, 3, device=dev)
    renderer.render_accumulate_device(nx, ny, 0, ns, full.data_ptr(), seed=77, variant=variant)
    parts = torch.zeros(ny, nx, 3, device=dev)
    renderer.render_accumulate_device(nx, ny, 0, 3, parts.data_ptr(), seed=77, variant=variant)   # samples [0,3)
    renderer.render_accumulate_device(nx, ny, 3, 5, parts.data_ptr(), seed=77, variant=variant)   # samples [3,8)
    rows = torch.zeros(ny, nx, 3, device=dev)
    for g in range(3):                                                                   # rows j = g (mod 3)
        renderer.render_accumulate_device(nx, ny, 0, ns, rows.data_ptr(), row_offset=g, row_stride=3, seed=77,
                                          variant=variant)
    torch.cuda.synchronize()
    f, p, r = full.cpu().numpy(), parts.cpu().numpy(), rows.cpu().numpy()
    assert np.isfinite(f).all() and np.isfinite(p).all() and np.isfinite(r).all()
    # same paths, different float summation order (atomics): equal up to rounding of ~8 addends
    for other in (p, r):
        err = np.abs(f - other) / np.maximum(1.0, np.abs(f))
        assert err.max() < 1e-5, (float(err.max()), np.unravel_index(err.argmax(), err.shape))
    assert f.sum() > 0
    rgb = torch.zeros(ny, nx, 3, dtype=torch.uint8, device=dev)
    renderer.resolve_device(nx, ny, ns, full.data_ptr(), rgb.data_ptr())
    torch.cuda.synchronize()
    assert np.array_equal(rgb.cpu().numpy(), oracle.resolve(f.astype(np.float64), ns))
    # NaN -> 0, clamp at 255 (core.clj:56)
    special = torch.tensor([[[float("nan"), 1e9, 0.25 * ns]]], device=dev)
    out = torch.zeros(1, 1, 3, dtype=torch.uint8, device=dev)
    renderer.resolve_device(1, 1, ns, special.data_ptr(), out.data_ptr())
    assert out.cpu().numpy().tolist() == [[[0, 255, 127]]]
    # rt_render (host buffers) returns the same image as the device path
    lin, img = renderer.render(nx, ny, ns, 50, seed=77, variant=variant)
    # (the megakernel inlines its shading code once per path slot; the copies round differently, so a
    # chaotic glass path may differ in the last bits between two runs: 1e-4, not 1e-5)
    assert np.allclose(lin * ns, f, rtol=1e-4, atol=1e-4)
    assert (img != rgb.cpu().numpy()).mean() < 1e-3     # float summation order may flip a rare 8-bit boundary


def test_error_paths(scene_c2):
    flat, cam_type, cam, _ = scene_c2
    with rt.native.Renderer([0]) as r:
        with pytest.raises(rt.native.NativeError, match=r"\(-4\)"):       # RT_ERR_STATE
            r.render(8, 8, 1)
        r.set_scene(flat)
        with pytest.raises(rt.native.NativeError, match=r"\(-4\)"):       # camera missing
            r.render(8, 8, 1)
        r.set_camera(cam_type, cam)
        with pytest.raises(rt.native.NativeError, match=r"\(-1\)"):       # RT_ERR_ARG
            r.render(0, 8, 1)
        bad = flat.copy()
        bad.material_id[3] = 10_000
        with pytest.raises(rt.native.NativeError, match=r"\(-1\)"):
            r.set_scene(bad)
        bad = flat.copy()
        bad.mat_type[0] = 9                                               # e.g. Isotropic: outside the path
        with pytest.raises(rt.native.NativeError, match=r"\(-2\)"):       # RT_ERR_UNSUPPORTED
            r.set_scene(bad)
        lin, img = r.render(16, 8, 2)                                     # the context is still usable
        assert lin.shape == (8, 16, 3) and img.shape == (8, 16, 3) and np.isfinite(lin).all()
    with pytest.raises(rt.native.NativeError):
        rt.native.Renderer([99])


def test_multi_device_in_library_matches_single(scene_c2):
    """rt_create over ALL visible devices (the single-process path a JVM caller uses, core.clj:99-108): sample slices or
    interleaved rows on each device, per-device float sums combined on device 0 — by NVLink peer loads inside the
    resolve kernel, or by one ncclReduce (libnccl.so.2 loaded on demand) — equal the single-device render."""
    import torch

    n_dev = torch.cuda.device_count()
    if n_dev < 2:
        pytest.skip("needs 2 GPUs")
    flat, cam_type, cam, _ = scene_c2
    nx, ny, ns = 160, 96, 16
    with rt.native.Renderer([0]) as r1:
        r1.set_scene(flat)
        r1.set_camera(cam_type, cam)
        a, img_a = r1.render(nx, ny, ns, 50, seed=11)
    for reduce_mode in (0, 1):
        for rows in (0, 1):
            with rt.native.Renderer(list(range(n_dev))) as r2:
                r2.set_option("reduce", reduce_mode)
                r2.set_option("rows", rows)
                r2.set_scene(flat)
                r2.set_camera(cam_type, cam)
                b, img_b = r2.render(nx, ny, ns, 50, seed=11)
                c2 = r2.counters()
            assert np.allclose(a, b, rtol=1e-4, atol=1e-4), (reduce_mode, rows)   # same paths; float summation order differs
            assert (img_a != img_b).mean() < 1e-3
            assert c2["samples"] == nx * ny * ns and c2["reduce_ns"] > 0
