"""GPU parity tests added in round 2 (all through the C ABI, against the CPU oracle).

  * DETERMINISTIC multi-bounce parity: rt_trace_paths runs the production kernels (both variants) for chosen
    (pixel, sample) pairs; the oracle replays the same paths on the same Philox counters (orc_trace_paths).
    Free-running comparison (bounce count, termination reason, radiance) on >= 1 M paths of BASELINE's own frames,
    plus a teacher-forced check of every logged bounce (hit id / t bit-exact, scattered ray within FP32 tolerance),
    which chaos cannot blur.
  * the device's closed-form ball / disk samplers against the reference's rejection sampling (util.clj:32-52), KS.
  * a full 1200x800x10 spp frame (BASELINE config 2) and a 256x256 crop of the 3840x2160 frame at 1024 spp
    (config 3) through the Monte-Carlo RMSE bound.
  * rectangles / triangles / wrappers / boxes / media / Perlin / image textures (hitable.clj:269-581,
    texture.clj:60-138): closest hit id equal and t within 1e-5 (bit-exact for everything but media, whose
    log() differs in the last ulp), shading within FP32 tolerance, renders inside the RMSE bound, both variants.
"""
import math
import random

import numpy as np
import pytest

import oracle
import raytrace_clj_b200 as rt
from raytrace_clj_b200 import hitable as hit
from raytrace_clj_b200 import shader as shad
from raytrace_clj_b200 import texture as tex
from raytrace_clj_b200.util import vec3

from helpers import RNG_DOMAIN, philox4x32_10, u01

pytestmark = pytest.mark.gpu

FMAX = float(np.finfo(np.float32).max)


@pytest.fixture(scope="module")
def renderer():
    r = rt.native.Renderer([0])
    yield r
    r.close()


def _rmse(a, b):
    return float(np.sqrt(np.mean((np.asarray(a, np.float64) - np.asarray(b, np.float64)) ** 2)))


def _load(renderer, sc):
    flat = rt.native.marshal_world(sc["world"])
    cam_type, cam = rt.native.marshal_camera(sc["camera"])
    renderer.set_scene(flat)
    renderer.set_camera(cam_type, cam)
    return flat, cam_type, cam, oracle.Scene(flat)


def _compare_paths(g, o, label, min_same=0.999, min_close=0.999):
    """g, o = (radiance, nrays, term, log) of the GPU and of the oracle replay."""
    same = (g[1] == o[1]) & (g[2] == o[2])
    ref = o[0]
    err = np.abs(g[0].astype(np.float64) - ref).max(axis=1)
    scale = np.maximum(np.abs(ref).max(axis=1), 1e-2)
    close = same & (err <= 1e-3 * scale)
    print(f"[paths] {label}: {len(same)} paths, same (bounces, termination) {same.mean():.5f}, "
          f"of those radiance within 1e-3 rel {close.sum() / max(1, same.sum()):.5f}; mean rays/path gpu {g[1].mean():.3f} "
          f"oracle {o[1].mean():.3f}; mean radiance gpu {g[0].astype(np.float64).mean():.6f} oracle {ref.mean():.6f}")
    assert same.mean() >= min_same, (label, same.mean())
    assert close.sum() / max(1, same.sum()) >= min_close, (label, close.sum() / same.sum())
    # the paths that diverged (a hit / miss or a Schlick coin decided differently in FP32) are unbiased
    dm = np.abs((g[0].astype(np.float64) - ref).mean(axis=0))
    assert np.all(dm < 5 * ref.std(axis=0) / math.sqrt(len(ref)) + 1e-5), dm
    return same


@pytest.mark.parametrize("variant", [1, 0])
@pytest.mark.parametrize("workload", ["c2", "c4"])
def test_trace_paths_replay_on_baseline_frames(renderer, workload, variant):
    """>= 1 M sampled paths of BASELINE config 2 (1200x800 random spheres) / config 4 (metal / glass heavy), depth 50:
    the production kernels and the oracle replay agree path for path."""
    nx, ny = 1200, 800
    if workload == "c2":
        sc, spp = rt.scene.make_random_scene(nx, ny, 11, True, random.Random(1)), 10
    else:
        sc, spp = rt.scene.make_material_stress_scene(nx, ny, 11, random.Random(4)), 64
    flat, cam_type, cam, S = _load(renderer, sc)
    n = 1_048_576 if variant == 1 else 262_144
    g = np.random.default_rng(17 + variant)
    pix = g.integers(0, nx * ny, n).astype(np.int32)
    smp = g.integers(0, spp, n).astype(np.int32)
    got = renderer.trace_paths(nx, ny, pix, smp, 50, seed=7, variant=variant)
    ref = S.trace_paths(cam_type, cam, nx, ny, pix, smp, 50, seed=7)
    # C4's long specular chains amplify FP32 rounding bounce after bounce: the free-running agreement is lower there
    # (the teacher-forced test below checks every bounce on its own)
    _compare_paths(got, ref, f"{workload} variant {variant}", min_same=0.999 if workload == "c2" else 0.99,
                   min_close=0.999 if workload == "c2" else 0.99)
    assert got[1].min() >= 1 and got[1].max() <= 51 and set(np.unique(got[2])) <= {1, 2, 3, 4}


def _ball_from_philox(seed, pixel, sample, bounce):
    """the device's rand_in_unit_sphere / scatter rand of (pixel, sample, bounce): block 1 (rt_device.cuh)."""
    n = len(pixel)
    ctr = np.stack([np.asarray(pixel, np.uint32), np.asarray(sample, np.uint32),
                    (np.asarray(bounce, np.uint32) << np.uint32(16)) | np.uint32(1), np.full(n, RNG_DOMAIN, np.uint32)], axis=1)
    key = np.tile(np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], np.uint32), (n, 1))
    b = philox4x32_10(ctr, key)
    ux, uy, uz, uw = (u01(b[:, k]).astype(np.float64) for k in range(4))
    rad, z = np.cbrt(ux), 1.0 - 2.0 * uy
    s = np.sqrt(np.maximum(0.0, 1.0 - z * z)) * rad
    phi = 2.0 * np.pi * uz
    return np.stack([s * np.cos(phi), s * np.sin(phi), z * rad], axis=1).astype(np.float32), uw.astype(np.float32)


@pytest.mark.parametrize("variant", [1, 0])
def test_bounce_log_teacher_forced(renderer, variant):
    """Every logged bounce of the production kernels checked ON ITS OWN (no error accumulates along the path): the
    oracle's hit? on the logged ray gives the logged (id, t) bit for bit; the oracle's scatter fed the same Philox
    draws gives the next logged ray within FP32 shading tolerance; the path ends where the oracle ends it."""
    nx, ny = 1200, 800
    sc = rt.scene.make_material_stress_scene(nx, ny, 11, random.Random(4))
    flat, cam_type, cam, S = _load(renderer, sc)
    n, LB = 150_000, 12
    g = np.random.default_rng(23)
    pix = g.integers(0, nx * ny, n).astype(np.int32)
    smp = g.integers(0, 64, n).astype(np.int32)
    rad, nr, term, log = renderer.trace_paths(nx, ny, pix, smp, 50, seed=11, variant=variant, log_bounces=LB)
    valid = log["hit_id"] != -2
    assert np.array_equal(valid.sum(axis=1), np.minimum(nr, LB))               # exactly the rays the path traced
    q, b = np.nonzero(valid)
    o, d, tm = log["o"][q, b], log["d"][q, b], log["time"][q, b]
    t_ref, id_ref = S.hit(o, d, tm, 0.001, FMAX)
    assert np.array_equal(id_ref, log["hit_id"][q, b])
    hitm = id_ref >= 0
    assert np.array_equal(t_ref[hitm], log["t"][q, b][hitm]) and np.all(np.isinf(log["t"][q, b][~hitm]))
    # first logged ray = get-ray of (pixel, sample) (camera.clj:35-48)
    go, gd, gt, _ = renderer.generate_rays(nx, ny, np.stack([pix % nx, pix // nx], axis=1), smp, seed=11)
    assert np.array_equal(log["o"][:, 0], go) and np.array_equal(log["d"][:, 0], gd) and np.array_equal(log["time"][:, 0], gt)
    # scatter: bounces that have a successor in the log
    has_next = np.zeros_like(valid)
    has_next[:, :-1] = valid[:, 1:]
    q, b = np.nonzero(valid & has_next)
    ball, uw = _ball_from_philox(11, pix[q], smp[q], b + 1)
    ref = S.shade_batch(log["o"][q, b], log["d"][q, b], log["time"][q, b], log["hit_id"][q, b], ball, uw)
    no, nd = log["o"][q, b + 1], log["d"][q, b + 1]
    cont = ref["flags"] == 1
    # a continuing path the oracle would have ended (grazing metal reflection) or a flipped Schlick coin: FP32 threshold cases
    scale = np.maximum(1.0, np.linalg.norm(ref["dir"], axis=1))
    good = cont & (np.abs(nd - ref["dir"]).max(axis=1) <= 3e-4 * scale) & (np.abs(no - ref["origin"]).max(axis=1) <= 3e-4)
    print(f"[teacher-forced] variant {variant}: {len(q)} scatters checked, {1 - good.mean():.2e} threshold flips")
    assert good.mean() > 0.9995
    # the last ray of each finished path: the oracle ends it the same way
    fin = np.nonzero(nr <= LB)[0]
    lb = nr[fin] - 1
    lo, ld, lt, lid = log["o"][fin, lb], log["d"][fin, lb], log["time"][fin, lb], log["hit_id"][fin, lb]
    assert np.array_equal(lid < 0, term[fin] == rt.native.RT_TERM_MISS)
    ball, uw = _ball_from_philox(11, pix[fin], smp[fin], lb + 1)
    ref = S.shade_batch(lo, ld, lt, lid, ball, uw)
    ended = (ref["flags"] == 0) | (lid < 0) | (nr[fin] == 51)
    assert ended.mean() > 0.9995
    types = np.where(lid >= 0, flat.mat_type[flat.material_id[np.maximum(lid, 0)]], -1)
    assert np.all(types[term[fin] == rt.native.RT_TERM_LIGHT] == rt.native.RT_MAT_DIFFUSE_LIGHT)
    absorbed = term[fin] == rt.native.RT_TERM_ABSORB
    assert absorbed.sum() > 100 and np.all(types[absorbed] == rt.native.RT_MAT_METAL)


def _ks(a, b):
    a, b = np.sort(a), np.sort(b)
    allv = np.concatenate([a, b])
    return float(np.abs(np.searchsorted(a, allv, side="right") / len(a) - np.searchsorted(b, allv, side="right") / len(b)).max())


def test_device_samplers_match_rejection_sampling(renderer):
    """rand_in_unit_sphere / rand_in_unit_disk of rt_device.cuh (closed form, one Philox block) against the reference's
    rejection loops (util.clj:32-52) run by the oracle: KS on radius, cos(theta), phi and the Cartesian marginals; and
    point for point against the oracle's closed-form map on the same counters."""
    n = 400_000
    crit = 1.95 * math.sqrt(2.0 / n)
    dev = renderer.sample_device(0, n, seed=3).astype(np.float64)
    rej = oracle.sample_ball(n, seed=99, replay=False)
    assert np.linalg.norm(dev, axis=1).max() < 1.0
    for f in (lambda p: np.linalg.norm(p, axis=1), lambda p: p[:, 2] / np.linalg.norm(p, axis=1),
              lambda p: np.arctan2(p[:, 1], p[:, 0]), lambda p: p[:, 0], lambda p: p[:, 1], lambda p: p[:, 2]):
        assert _ks(f(dev), f(rej)) < crit
    assert np.abs(dev - oracle.sample_ball(n, seed=3, replay=True)).max() < 2e-6
    dev = renderer.sample_device(1, n, seed=5).astype(np.float64)
    rej = oracle.sample_disk(n, seed=77, replay=False)
    assert np.linalg.norm(dev, axis=1).max() < 1.0
    for f in (lambda p: np.linalg.norm(p, axis=1), lambda p: np.arctan2(p[:, 1], p[:, 0]), lambda p: p[:, 0], lambda p: p[:, 1]):
        assert _ks(f(dev), f(rej)) < crit
    assert np.abs(dev - oracle.sample_disk(n, seed=5, replay=True)).max() < 2e-6


def _rmse_bound_check(g1, g2, o1, o2, label, npix):
    g1, g2, o1, o2 = (np.asarray(x, np.float64) for x in (g1, g2, o1, o2))   # float32 sums over ~1e6 pixels lose 1e-3
    r_gg, r_oo = _rmse(g1, g2), _rmse(o1, o2)
    cross = [_rmse(g, o) for g in (g1, g2) for o in (o1, o2)]
    bound = 1.25 * math.sqrt((r_gg ** 2 + r_oo ** 2) / 2)
    print(f"[parity] {label}: RMSE cross {max(cross):.4f} <= bound {bound:.4f} (R_gg {r_gg:.4f}, R_oo {r_oo:.4f})")
    assert max(cross) <= bound, (cross, r_gg, r_oo)
    assert 0.8 < r_gg / r_oo < 1.25
    ax = tuple(range(g1.ndim - 1))
    sigma = np.sqrt(((o1 - o2) ** 2).mean(axis=ax) / 2 + ((g1 - g2) ** 2).mean(axis=ax) / 2)
    mean_diff = np.abs((g1 + g2).mean(axis=ax) / 2 - (o1 + o2).mean(axis=ax) / 2)
    assert np.all(mean_diff <= 3 * sigma / math.sqrt(npix) + 2e-4), (mean_diff, sigma)


def test_full_c2_frame_matches_oracle(renderer):
    """BASELINE config 2 at its own size: 1200x800, 10 spp, depth 50 (9.6 M paths per render), wavefront variant."""
    nx, ny, ns = 1200, 800, 10
    flat, cam_type, cam, S = _load(renderer, rt.scene.make_random_scene(nx, ny, 11, True, random.Random(1)))
    g1, _ = renderer.render(nx, ny, ns, 50, seed=101, rgb8=False)
    g2, _ = renderer.render(nx, ny, ns, 50, seed=202, rgb8=False)
    o1 = S.render_accumulate(cam_type, cam, nx, ny, 0, ns, 50, seed=303)[0] / ns
    o2 = S.render_accumulate(cam_type, cam, nx, ny, 0, ns, 50, seed=404)[0] / ns
    _rmse_bound_check(g1, g2, o1, o2, "c2 full frame 1200x800x10", nx * ny)
    # and the SAME frame sample for sample: the oracle replaying the GPU's Philox counters (seed 101)
    o3 = S.render_accumulate(cam_type, cam, nx, ny, 0, ns, 50, seed=101, replay=True)[0] / ns
    d = np.abs(g1.astype(np.float64) - o3).max(axis=2)
    print(f"[parity] c2 full frame, replay of the same seed: pixels within 1e-3: {(d <= 1e-3).mean():.5f}, RMSE {_rmse(g1, o3):.5f}")
    assert (d <= 1e-3).mean() > 0.99 and _rmse(g1, o3) < 0.25 * _rmse(g1, o1)


def test_c3_crop_1024spp_matches_oracle(renderer):
    """BASELINE config 3's frame (3840x2160, 1024 spp): a 256x256 crop around the three hero spheres, all 1024 samples
    of every pixel of the crop, through rt_trace_paths (the production kernels) vs the oracle, RMSE bound."""
    nx, ny, ns, C = 3840, 2160, 1024, 256
    flat, cam_type, cam, S = _load(renderer, rt.scene.make_random_scene(nx, ny, 11, True, random.Random(1)))
    i0, j0 = nx // 2 - C // 2, ny // 2 - C // 2 - 100
    jj, ii = np.meshgrid(np.arange(j0, j0 + C), np.arange(i0, i0 + C), indexing="ij")
    crop_pix = (jj * nx + ii).reshape(-1).astype(np.int32)

    def crop(fn, seed):
        acc = np.zeros((C * C, 3), np.float64)
        step = 64
        for s0 in range(0, ns, step):
            pix = np.repeat(crop_pix, step)
            smp = np.tile(np.arange(s0, s0 + step, dtype=np.int32), C * C)
            rad = fn(pix, smp, seed)
            acc += np.asarray(rad, np.float64).reshape(C * C, step, 3).sum(axis=1)
        return (acc / ns).reshape(C, C, 3)

    gpu = lambda pix, smp, seed: renderer.trace_paths(nx, ny, pix, smp, 50, seed=seed)[0]          # noqa: E731
    cpu = lambda pix, smp, seed: S.trace_paths(cam_type, cam, nx, ny, pix, smp, 50, seed=seed)[0]   # noqa: E731
    g1, g2 = crop(gpu, 1), crop(gpu, 2)
    o1, o2 = crop(cpu, 3), crop(cpu, 4)
    _rmse_bound_check(g1, g2, o1, o2, "c3 256x256 crop at 1024 spp", C * C)


# ---- f-2: rectangles, triangles, boxes, wrappers ------------------------------------------------------------------
def _assert_hits(t_gpu, id_gpu, t_ref, id_ref, exact=True):
    assert np.array_equal(id_gpu, id_ref), f"{int((id_gpu != id_ref).sum())} of {len(id_ref)} ids differ"
    h = id_ref >= 0
    assert np.all(np.isinf(t_gpu[~h]))
    rel = np.abs(t_gpu[h] - t_ref[h]) / np.abs(t_ref[h])
    assert rel.max(initial=0.0) <= 1e-5
    if exact:
        assert np.array_equal(t_gpu[h], t_ref[h])


def _rays_at(flat, S, g, n, origin_box, jitter=1e-6):
    """n rays from random origins aimed at random points of random leaves' bounding boxes — including their edges and
    corners (+- jitter), where the inclusive / strict tests decide."""
    k = g.integers(0, flat.n_spheres, n)
    bb = np.array([S.prim_bbox(i) for i in range(flat.n_spheres)])
    lo, hi = bb[k, :3], bb[k, 3:]
    w = g.random((n, 3))
    snap = g.random((n, 3))
    w = np.where(snap < 0.15, 0.0, np.where(snap > 0.85, 1.0, w))       # 30 % of the coordinates sit ON a box face / edge
    target = lo + w * (hi - lo) + g.normal(scale=jitter, size=(n, 3)) * (hi - lo + 1e-3)
    o = g.uniform(origin_box[0], origin_box[1], size=(n, 3))
    d = (target - o) * g.uniform(0.3, 3.0, size=(n, 1))
    return o.astype(np.float32), d.astype(np.float32), g.random(n).astype(np.float32)


@pytest.mark.parametrize("name", ["cornell", "triangles", "light", "boxes"])
def test_trace_primary_new_primitives(renderer, name):
    """id equal and t bit-identical to the oracle on >= 200 k rays per scene, incl. edge-on and edge-grazing rays."""
    rng = random.Random(3)
    if name == "cornell":
        sc, box = rt.scene.make_cornell_box(100, 100, True, rng), ((-100, -100, -900), (655, 655, 655))
    elif name == "triangles":
        sc, box = rt.scene.make_two_triangles(100, 100, rng), ((-3, -3, -12), (5, 5, 12))
    elif name == "light":
        sc, box = rt.scene.make_example_light(100, 100, rng), ((-15, -2, -15), (15, 12, 15))
    else:   # many small rotated / translated boxes + triangles in a Hitlist: hundreds of generic leaves through the cull
        items = []
        m = shad.lambertian(albedo=tex.constant(color=vec3(.7, .7, .7)))
        for _ in range(60):
            p1 = vec3(rng.uniform(.2, 1.5), rng.uniform(.2, 1.5), rng.uniform(.2, 1.5))
            b = hit.box(p0=vec3(0, 0, 0), p1=p1, material=m)
            items.append(hit.translate(item=hit.rotate_y(item=b, theta=rng.uniform(-90, 90)),
                                       offset=vec3(rng.uniform(-10, 10), rng.uniform(0, 3), rng.uniform(-10, 10))))
        for _ in range(60):
            v0 = vec3(rng.uniform(-10, 10), rng.uniform(0, 4), rng.uniform(-10, 10))
            items.append(hit.triangle(v0=v0, v1=v0 + vec3(rng.uniform(-1, 1), rng.uniform(.2, 1), rng.uniform(-1, 1)),
                                      v2=v0 + vec3(rng.uniform(.2, 1), rng.uniform(-1, 1), rng.uniform(-1, 1)), material=m))
        items.append(hit.sphere(center=vec3(0, -1000, 0), radius=1000, material=m))
        sc = {"world": hit.hitlist(items=items), "camera": rt.scene.make_two_spheres(100, 100)["camera"]}
        box = ((-14, -1, -14), (14, 8, 14))
    flat, cam_type, cam, S = _load(renderer, sc)
    g = np.random.default_rng(5)
    n = 220_000
    o, d, tm = _rays_at(flat, S, g, n, box)
    t_ref, id_ref = S.hit(o, d, tm, 0.001, FMAX)
    t_gpu, id_gpu = renderer.trace_primary(o, d, tm, 0.001, FMAX)
    _assert_hits(t_gpu, id_gpu, t_ref, id_ref)
    assert (id_ref >= 0).mean() > 0.3
    kinds = set(flat.prim_type[np.unique(id_ref[id_ref >= 0])].tolist())
    assert kinds >= {"cornell": {1, 2, 3}, "triangles": {0, 4}, "light": {0, 1}, "boxes": {0, 1, 2, 3, 4}}[name]
    # edge-on: rays lying IN a rectangle's plane (t = 0/0 or +-inf: the reference's comparisons reject them)
    if name == "cornell":
        oo = np.array([[278, 0, -800], [0, 278, -800], [278, 555, -800]], np.float32)
        dd = np.array([[0, 0, 1], [0, 0, 1], [0.1, 0, 1]], np.float32)
        _assert_hits(*renderer.trace_primary(oo, dd, None, 0.001, FMAX), *S.hit(oo, dd, None, 0.001, FMAX))
    # secondary rays from the hit points (self-intersection at t_min = 0.001 on flat primitives)
    h = id_ref >= 0
    p = (o[h].astype(np.float64) + t_ref[h][:, None] * d[h].astype(np.float64)).astype(np.float32)
    nd = g.normal(size=p.shape).astype(np.float32)
    _assert_hits(*renderer.trace_primary(p, nd, tm[h], 0.001, FMAX), *S.hit(p, nd, tm[h], 0.001, FMAX))
    # the FP32 cull (bounding spheres of the leaves) never loses a pair the exact test accepts
    lost, surv, cand = renderer.cull_check(o[:60000], d[:60000], tm[:60000], 0.001, FMAX)
    assert lost == 0 and surv >= cand


def test_shade_batch_new_primitives(renderer):
    """hit records of rectangles / triangles / wrapped boxes (p, normal through RotateY / FlipNormals, uv) feeding the
    scatter functions: scattered ray, attenuation, emitted vs the oracle with caller-given randoms."""
    flat, cam_type, cam, S = _load(renderer, rt.scene.make_cornell_box(100, 100, True, random.Random(3)))
    g = np.random.default_rng(6)
    o, d, tm = _rays_at(flat, S, g, 60_000, ((100, 100, -800), (455, 455, 500)), jitter=0)
    t, ids = S.hit(o, d, tm)
    keep = ids >= 0
    o, d, tm, t, ids = o[keep], d[keep], tm[keep], t[keep], ids[keep]
    ball = g.uniform(-1, 1, size=(len(o), 3))
    ball *= (g.random(len(o)) ** (1 / 3) / np.linalg.norm(ball, axis=1))[:, None]
    u = g.random(len(o)).astype(np.float32)
    got = renderer.shade_batch(o, d, tm, ids, t, ball.astype(np.float32), u)
    ref = S.shade_batch(o, d, tm, ids, ball.astype(np.float32), u)
    assert np.array_equal(got["flags"], ref["flags"])
    c = ref["flags"] == 1
    assert c.sum() > 10_000 and (~c).sum() > 50                   # Lambertian walls scatter, the light does not
    assert np.allclose(got["dir"][c], ref["dir"][c], atol=2e-4)
    assert np.allclose(got["origin"][c], ref["origin"][c], rtol=1e-6, atol=2e-3)   # coordinates up to 555
    assert np.allclose(got["atten"][c], ref["atten"][c], atol=1e-6) and np.allclose(got["emitted"], ref["emitted"], atol=1e-6)
    assert len(np.unique(ids)) >= 15                             # walls, light and both rotated blocks' faces


def _render_parity(renderer, sc, nx, ny, ns, variant, label, depth=50):
    flat, cam_type, cam, S = _load(renderer, sc)
    g1, img = renderer.render(nx, ny, ns, depth, seed=101, variant=variant)
    g2, _ = renderer.render(nx, ny, ns, depth, seed=202, variant=variant)
    o1 = S.render_accumulate(cam_type, cam, nx, ny, 0, ns, depth, seed=303)[0] / ns
    o2 = S.render_accumulate(cam_type, cam, nx, ny, 0, ns, depth, seed=404)[0] / ns
    assert np.isfinite(g1).all()
    _rmse_bound_check(g1, g2, o1, o2, f"{label} {nx}x{ny}x{ns} variant {variant}", nx * ny)
    return flat, cam_type, cam, S


@pytest.mark.parametrize("variant", [1, 0])
@pytest.mark.parametrize("name", ["cornell", "triangles", "light"])
def test_render_new_primitive_scenes(renderer, name, variant):
    """make-cornell-box classic (scene.clj:230), make-two-triangles (:80), make-example-light (:191)."""
    rng = random.Random(2)
    if name == "cornell":
        _render_parity(renderer, rt.scene.make_cornell_box(96, 96, True, rng), 96, 96, 256, variant, "cornell box")
    elif name == "triangles":
        _render_parity(renderer, rt.scene.make_two_triangles(120, 80, rng), 120, 80, 64, variant, "two triangles")
    else:
        _render_parity(renderer, rt.scene.make_example_light(120, 80, rng), 120, 80, 256, variant, "example light")


@pytest.mark.parametrize("variant", [1, 0])
def test_cornell_paths_replay(renderer, variant):
    """Deterministic path replay on the Cornell box (long diffuse paths between rectangles and rotated boxes)."""
    nx = ny = 200
    flat, cam_type, cam, S = _load(renderer, rt.scene.make_cornell_box(nx, ny, True, random.Random(2)))
    g = np.random.default_rng(31)
    n = 200_000
    pix = g.integers(0, nx * ny, n).astype(np.int32)
    smp = g.integers(0, 256, n).astype(np.int32)
    got = renderer.trace_paths(nx, ny, pix, smp, 50, seed=5, variant=variant)
    ref = S.trace_paths(cam_type, cam, nx, ny, pix, smp, 50, seed=5)
    _compare_paths(got, ref, f"cornell variant {variant}", min_same=0.995, min_close=0.995)


# ---- f-4: media, Perlin, image maps ------------------------------------------------------------------------------------
@pytest.mark.parametrize("variant", [1, 0])
@pytest.mark.parametrize("name", ["perlin", "earth", "subsurface", "smoke", "final"])
def test_render_volume_and_texture_scenes(renderer, name, variant):
    rng = random.Random(2)
    if name == "perlin":
        _render_parity(renderer, rt.scene.make_two_perlin_spheres(120, 80, rng), 120, 80, 64, variant, "two perlin spheres")
    elif name == "earth":
        _render_parity(renderer, rt.scene.make_textured_sphere(120, 80, rng), 120, 80, 64, variant, "textured sphere")
    elif name == "subsurface":
        _render_parity(renderer, rt.scene.make_subsurface_sphere(120, 80, rng), 120, 80, 128, variant, "subsurface sphere")
    elif name == "smoke":
        _render_parity(renderer, rt.scene.make_cornell_box(80, 80, False, rng), 80, 80, 256, variant, "cornell smoke")
    else:
        _render_parity(renderer, rt.scene.make_final(96, 96, rng, nb=8, ns=200), 96, 96, 256, variant, "final (book 2)")


def test_texture_samples_on_device(renderer):
    """Perlin turbulence / marble / flipped image map evaluated by the shading kernel vs the oracle (FP32 noise)."""
    img = rt.scene.synthetic_earth(64, 32)
    mats = [shad.lambertian(albedo=tex.perlin_turbulence(scale=4, depth=7)),
            shad.lambertian(albedo=tex.marble(scale=0.1, depth=4)),
            shad.lambertian(albedo=tex.perlin_noise(scale=3)),
            shad.lambertian(albedo=tex.flip_texture_v(tex=tex.image_map(image=img)))]
    items = [hit.uv_sphere(center=vec3(20.0 * k, 0, 0), radius=1, material=m) for k, m in enumerate(mats)]
    flat = rt.native.marshal_world(hit.hitlist(items=items))
    renderer.set_scene(flat)
    S = oracle.Scene(flat)
    g = np.random.default_rng(8)
    n = 40_000
    k = g.integers(0, 4, n)
    dirs = g.normal(size=(n, 3))
    dirs /= np.linalg.norm(dirs, axis=1)[:, None]
    o = (np.stack([20.0 * k, 0 * k, 0 * k], axis=1) + 4.0 * dirs).astype(np.float32)
    d = (-dirs).astype(np.float32)
    tm = np.zeros(n, np.float32)
    t, ids = S.hit(o, d, tm)
    assert np.array_equal(ids, k)
    ball = np.zeros((n, 3), np.float32)
    got = renderer.shade_batch(o, d, tm, ids, t, ball, np.zeros(n, np.float32))
    ref = S.shade_batch(o, d, tm, ids, ball, np.zeros(n, np.float32))
    err = np.abs(got["atten"] - ref["atten"]).max(axis=1)
    for kk, tol in ((0, 5e-3), (1, 5e-3), (2, 1e-4)):
        assert err[k == kk].max() < tol, (kk, err[k == kk].max())
        assert ref["atten"][k == kk].std() > 0.02                 # the texture really varies
    # image map: a texel boundary may fall between the FP32 and the double uv: all but a sliver agree exactly
    assert (err[k == 3] < 1e-6).mean() > 0.995


def test_medium_hits_on_device(renderer):
    """ConstantMedium.hit? (hitable.clj:516-543) in the FP64 refine: a fog ball and a fog box (rotated, translated),
    rays from outside, from inside, grazing; the medium's `rand` is 0.5 without a path (as in the oracle)."""
    ball = hit.sphere(center=vec3(0, 0, 0), radius=2, material=shad.dielectric(ri=1.5))
    blk = hit.translate(item=hit.rotate_y(item=hit.box(p0=vec3(0, 0, 0), p1=vec3(2, 3, 2), material=None), theta=30.0),
                        offset=vec3(5, -1, 0))
    world = hit.hitlist(items=[hit.constant_medium(boundary=ball, density=0.4, albedo=tex.constant(color=vec3(.2, .4, .9))),
                               hit.constant_medium(boundary=blk, density=0.7, albedo=tex.constant(color=vec3(1, 1, 1)))])
    flat = rt.native.marshal_world(world)
    renderer.set_scene(flat)
    S = oracle.Scene(flat)
    g = np.random.default_rng(4)
    n = 100_000
    o = g.uniform(-6, 10, size=(n, 3)).astype(np.float32)
    o[: n // 4] = (g.normal(size=(n // 4, 3)) * 0.8).astype(np.float32)                 # inside the ball
    target = np.where(g.random((n, 1)) < 0.5, [0, 0, 0], [6, 0.5, 1]) + g.normal(scale=1.0, size=(n, 3))
    d = ((target - o) * g.uniform(0.2, 2.0, size=(n, 1))).astype(np.float32)
    t_ref, id_ref = S.hit(o, d, None, 0.001, FMAX)
    t_gpu, id_gpu = renderer.trace_primary(o, d, None, 0.001, FMAX)
    assert (id_ref == 0).sum() > 5000 and (id_ref == 1).sum() > 5000 and (id_ref < 0).sum() > 5000
    _assert_hits(t_gpu, id_gpu, t_ref, id_ref, exact=False)        # log() may differ in the last ulp: 1e-5 relative, not bit-exact
    h = id_ref >= 0
    assert (t_gpu[h] == t_ref[h]).mean() > 0.9


# ---- f-3: the GPU BVH accelerator -----------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["random", "sweep6000", "cornell", "final"])
def test_bvh_accel_equals_brute_force(renderer, name):
    """RT_ACCEL_BVH (flattened tree, conservative FP32 slab tests, the same FP64 leaf tests and tie keys) returns the
    brute-force closest hit bit for bit — ids and t on camera / secondary / grazing rays, and whole paths: per-path
    radiance, bounce count and termination of rt_trace_paths are IDENTICAL under both accelerators."""
    rng = random.Random(2)
    nx, ny = 320, 200
    sc = {"random": lambda: rt.scene.make_random_scene(nx, ny, 11, True, rng),
          "sweep6000": lambda: rt.scene.make_scale_sweep_scene(nx, ny, 6000, rng),
          "cornell": lambda: rt.scene.make_cornell_box(nx, ny, True, rng),
          "final": lambda: rt.scene.make_final(nx, ny, rng, nb=10, ns=300)}[name]()
    flat, cam_type, cam, S = _load(renderer, sc)
    g = np.random.default_rng(9)
    n = 120_000
    camd = np.asarray(cam, np.float64)
    o = np.tile(camd[0:3], (n, 1)).astype(np.float32)
    d = (camd[3:6] + g.random((n, 1)) * camd[6:9] + g.random((n, 1)) * camd[9:12] - camd[0:3]).astype(np.float32)
    tm = g.random(n).astype(np.float32)
    try:
        renderer.set_accel(rt.native.RT_ACCEL_BRUTE_FORCE)
        t0, id0 = renderer.trace_primary(o, d, tm)
        renderer.set_accel(rt.native.RT_ACCEL_BVH)
        t1, id1 = renderer.trace_primary(o, d, tm)
        assert np.array_equal(id0, id1) and np.array_equal(t0, t1)
        t_ref, id_ref = S.hit(o, d, tm)
        exact = name != "final"                               # a medium's log() may differ in the last ulp from the CPU's
        _assert_hits(t1, id1, t_ref, id_ref, exact=exact)
        h = id0 >= 0
        p = (o[h].astype(np.float64) + t0[h][:, None] * d[h].astype(np.float64)).astype(np.float32)
        nd = g.normal(size=p.shape).astype(np.float32)
        a = renderer.trace_primary(p, nd, tm[h])
        renderer.set_accel(rt.native.RT_ACCEL_BRUTE_FORCE)
        b = renderer.trace_primary(p, nd, tm[h])
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
        # whole paths, both accelerators, the production wavefront
        m = 150_000
        pix = g.integers(0, nx * ny, m).astype(np.int32)
        smp = g.integers(0, 64, m).astype(np.int32)
        brute = renderer.trace_paths(nx, ny, pix, smp, 50, seed=3)
        renderer.reset_counters()
        renderer.set_accel(rt.native.RT_ACCEL_BVH)
        tree = renderer.trace_paths(nx, ny, pix, smp, 50, seed=3)
        c = renderer.counters()
        brute2 = None
        if not (np.array_equal(brute[1], tree[1]) and np.array_equal(brute[0], tree[0])):      # diagnostics before failing
            renderer.set_accel(rt.native.RT_ACCEL_BRUTE_FORCE)
            brute2 = renderer.trace_paths(nx, ny, pix, smp, 50, seed=3, log_bounces=6)
            renderer.set_accel(rt.native.RT_ACCEL_BVH)
            tree2 = renderer.trace_paths(nx, ny, pix, smp, 50, seed=3, log_bounces=6)
            bad = np.nonzero((brute[1] != tree[1]) | np.any(brute[0] != tree[0], axis=1))[0]
            print(f"[bvh] {name}: {len(bad)} of {m} paths differ; brute self-consistent: {np.array_equal(brute[0], brute2[0])}; "
                  f"tree self-consistent: {np.array_equal(tree[0], tree2[0])}")
            for q in bad[:6]:
                print("   path", q, "nrays", brute[1][q], tree[1][q], "brute hits", brute2[3]["hit_id"][q], brute2[3]["t"][q],
                      "tree hits", tree2[3]["hit_id"][q], tree2[3]["t"][q], "times", brute2[3]["time"][q])
        assert np.array_equal(brute[1], tree[1]) and np.array_equal(brute[2], tree[2]) and np.array_equal(brute[0], tree[0])
        assert c["bvh_node_tests"] > 0 and c["rays"] == int(tree[1].sum())
        if name in ("random", "sweep6000"):                   # O(log N): far fewer exact tests than rays x N
            assert c["sphere_tests"] < 0.05 * c["rays"] * flat.n_spheres
        lin, _ = renderer.render(nx, ny, 8, 50, seed=5)
        renderer.set_accel(rt.native.RT_ACCEL_BRUTE_FORCE)
        lin0, _ = renderer.render(nx, ny, 8, 50, seed=5)
        assert np.allclose(lin, lin0, rtol=1e-5, atol=1e-6)    # same paths; only the order of the float atomics differs
    finally:
        renderer.set_accel(rt.native.RT_ACCEL_BRUTE_FORCE)


# ------------------------------------------------------------------------------------------
# tensor-core cull (rt_cull_tc.cuh, option cull_tc): same contract as the FP32 cull, same results downstream
# ------------------------------------------------------------------------------------------
def _tc_ray_families(flat, cam, rng):
    from helpers import camera_rays
    fam = {}
    o, d, tm = camera_rays(cam, 1200, 800, 60_000, rng)
    fam["camera"] = (o, d, tm)
    o2 = rng.uniform(-15, 15, size=(60_000, 3)).astype(np.float32)
    o2[:, 1] = rng.uniform(-1, 3, size=len(o2))
    fam["volume"] = (o2, rng.normal(size=o2.shape).astype(np.float32), rng.random(len(o2)).astype(np.float32))
    for scale in (1.0, 30.0, 1000.0):
        for jitter in (3e-7, 1e-4):
            n = 30_000
            k = rng.integers(0, flat.n_spheres, n)
            c = flat.center0_r[k, :3].astype(np.float64)
            rad = np.abs(flat.center0_r[k, 3].astype(np.float64))
            oo = rng.normal(size=(n, 3)) * scale
            to_c = c - oo
            dist = np.linalg.norm(to_c, axis=1)
            perp = np.cross(to_c, rng.normal(size=(n, 3)))
            perp /= np.linalg.norm(perp, axis=1)[:, None]
            target = c + perp * (rad * (1.0 + rng.normal(scale=jitter, size=n)))[:, None]
            dd = (target - oo) * rng.uniform(0.2, 3.0, size=(n, 1)) / dist[:, None]
            fam[f"grazing |o|~{scale:g} +-{jitter:g}"] = (oo.astype(np.float32), dd.astype(np.float32), rng.random(n).astype(np.float32))
    return fam


def test_tensor_core_cull_never_under_reports(renderer, random_scene_flat):
    """wf_cull_tc — the PRODUCTION kernel, run over a queue of the caller's rays — must emit every (ray, leaf) pair the exact
    FP64 test accepts (it may over-report), and no more pairs than the FP32 cull it confirms its candidates with."""
    flat, cam_type, cam = random_scene_flat
    renderer.set_scene(flat)
    fam = _tc_ray_families(flat, cam, np.random.default_rng(77))
    try:
        for name, (o, d, tm) in fam.items():
            res = {}
            for mode in (0, 1):
                renderer.set_option("cull_tc", mode)
                res[mode] = renderer.cull_check(o, d, tm, 0.001, FMAX)
                lost, surv, cand = res[mode]
                assert lost == 0, f"{name} cull_tc={mode}: lost {lost} of {cand} exact candidates"
            assert res[0][2] == res[1][2]
            if "1000" not in name:     # (far origins overflow the pair buffer in this diagnostic: those entries count as all-kept)
                # (the confirm step is the FP32 cull's own key; its per-ray constants are compiled in another kernel, so a
                # borderline pair may round the other way: allow a few in 10^5)
                assert res[1][1] <= res[0][1] * (1 + 1e-4) + 2, f"{name}: tensor-core survivors {res[1][1]} > FP32 survivors {res[0][1]}"
    finally:
        renderer.set_option("cull_tc", 1)


@pytest.mark.parametrize("scene_name", ["random", "cornell", "stress"])
def test_tensor_core_cull_gives_identical_paths(renderer, scene_name):
    """The cull only feeds the exact FP64 tests, so whole paths (radiance, bounce count, termination) must be IDENTICAL
    with the cull on the tensor cores and on the FP32 pipe."""
    import bench
    nx, ny = 320, 200
    flat, cam_type, cam = bench.build_scene(scene_name, nx, ny, 1)
    renderer.set_scene(flat)
    renderer.set_camera(cam_type, cam)
    rng = np.random.default_rng(3)
    n = 200_000
    pix = rng.integers(0, nx * ny, n).astype(np.int32)
    smp = rng.integers(0, 64, n).astype(np.int32)
    out = {}
    try:
        for mode in (0, 1):
            renderer.set_option("cull_tc", mode)
            out[mode] = renderer.trace_paths(nx, ny, pix, smp, 50, seed=9)
    finally:
        renderer.set_option("cull_tc", 1)
    for a, b in zip(out[0][:3], out[1][:3]):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("scene_name", ["sweep:100", "cornell", "random"])
def test_tensor_core_cull_short_queues_and_partial_tiles(renderer, scene_name):
    """Launches of a few rays: one ray tile per CTA, the last one mostly DEAD slots, and (scenes with few leaves) feature
    tiles that are mostly padding rows.  Regression: dead ray x padding row used to come out as +0 = a candidate."""
    import bench
    flat, cam_type, cam = bench.build_scene(scene_name, 300, 300, 1)
    renderer.set_scene(flat)
    renderer.set_camera(cam_type, cam)
    rng = np.random.default_rng(12)
    try:
        for n in (1, 7, 128, 129, 130, 257, 1000, 4000):
            o = rng.uniform(-5, 5, size=(n, 3)).astype(np.float32) * (60.0 if scene_name == "cornell" else 1.0)
            d = rng.normal(size=(n, 3)).astype(np.float32)
            res = {}
            for mode in (0, 1):
                renderer.set_option("cull_tc", mode)
                res[mode] = renderer.cull_check(o, d, None, 0.001, FMAX)
                assert res[mode][0] == 0, f"{scene_name} n={n} cull_tc={mode}: lost {res[mode][0]} of {res[mode][2]}"
            assert res[0][2] == res[1][2]
    finally:
        renderer.set_option("cull_tc", 1)


def test_tensor_core_cull_long_lists_take_several_passes(renderer):
    """Over 1024 leaves the tensor-core cull covers the list in several launches per iteration (1024 leaves' features resident
    in shared memory each): same contract, same paths as the FP32 loop."""
    import bench
    nx, ny = 320, 180
    flat, cam_type, cam = bench.build_scene("sweep:3000", nx, ny, 5)
    renderer.set_scene(flat)
    renderer.set_camera(cam_type, cam)
    rng = np.random.default_rng(8)
    n = 20_000
    side = math.sqrt(flat.n_spheres) / 2
    o = rng.uniform(-side, side, size=(n, 3)).astype(np.float32)
    o[:, 1] = rng.uniform(0, 3, n)                       # in and just above the layer of spheres: dozens of leaves along many rays
    d = rng.normal(size=(n, 3)).astype(np.float32)
    pix = rng.integers(0, nx * ny, 100_000).astype(np.int32)
    smp = rng.integers(0, 16, 100_000).astype(np.int32)
    res, out = {}, {}
    try:
        for mode in (0, 1):
            renderer.set_option("cull_tc", mode)
            res[mode] = renderer.cull_check(o, d, None, 0.001, FMAX)
            assert res[mode][0] == 0
            out[mode] = renderer.trace_paths(nx, ny, pix, smp, 50, seed=9)
    finally:
        renderer.set_option("cull_tc", 1)
    assert res[0][2] == res[1][2]      # (these rays skim the layer: some overflow their warp's candidate list and count as all-kept)
    for a, b in zip(out[0][:3], out[1][:3]):
        assert np.array_equal(a, b)


TAIL_SOLO_DEFAULT = (24, 32)   # rt_api.cu Options::tail_solo / tail_lpp


@pytest.mark.parametrize("scene_name", ["random", "cornell", "final", "sweep:3000", "sweep:5000"])
def test_tail_warp_per_path_gives_identical_paths(renderer, scene_name):
    """wf_tail finishes the thin end of its slices one lane group per path (wf_solo_paths: leaves dealt over the group's lanes,
    closest hit by a shuffle minimum, shading replicated on the lanes).  Same cull key, same exact FP64 tests, same merge rule,
    same Philox blocks as the staged kernels: whole paths (radiance, bounce count, termination) and the bounce logs must be
    IDENTICAL whether a slice goes solo never (tail_solo = 0), at a few dozen paths, or from the tail's first bounce on
    (1 << 20), and whatever the group size (tail_lpp = 8 | 16 | 32 lanes per path).  (sweep:5000: over 4096 leaves the list is
    streamed through shared memory in tiles, and wf_solo_paths reads the cull records from global memory.)"""
    import bench
    nx, ny = 320, 200
    flat, cam_type, cam = bench.build_scene(scene_name, nx, ny, 1)
    renderer.set_scene(flat)
    renderer.set_camera(cam_type, cam)
    rng = np.random.default_rng(5)
    n = 120_000 if scene_name in ("random", "cornell") else 70_000
    pix = rng.integers(0, nx * ny, n).astype(np.int32)
    smp = rng.integers(0, 64, n).astype(np.int32)
    out, ctr = {}, {}
    modes = [(0, 32), (24, 32), (1 << 20, 32), (1 << 20, 8), (64, 16), (96, 8)]
    try:
        for solo, lpp in modes:
            renderer.set_option("tail_solo", solo)
            renderer.set_option("tail_lpp", lpp)
            renderer.reset_counters()
            out[solo, lpp] = renderer.trace_paths(nx, ny, pix, smp, 50, seed=21, log_bounces=6)
            ctr[solo, lpp] = renderer.counters()
    finally:
        renderer.set_option("tail_solo", TAIL_SOLO_DEFAULT[0])
        renderer.set_option("tail_lpp", TAIL_SOLO_DEFAULT[1])
    for mode in modes[1:]:
        for a, b in zip(out[modes[0]], out[mode]):
            if isinstance(a, np.ndarray):
                assert a.tobytes() == b.tobytes(), f"{scene_name}: (tail_solo, tail_lpp) = {mode} differs from the staged tail"
        for key in ("rays", "samples", "term_light", "term_absorb", "term_depth", "term_miss"):
            assert ctr[modes[0]][key] == ctr[mode][key], f"{scene_name}: counter {key} at {mode}: {ctr[mode][key]} != {ctr[modes[0]][key]}"


def test_render_into_page_locked_buffers(renderer, random_scene_flat):
    """rt_host_alloc: rt_render has the device write page-locked destinations directly (no staging copy); pageable ones go
    through the context's staging buffer.  Same frame either way (same seed: the float sums only differ by the order of the
    atomic adds), and a VIEW into the middle of a page-locked block is recognised too."""
    flat, cam_type, cam = random_scene_flat
    renderer.set_scene(flat)
    nx, ny = 240, 160
    sc = rt.scene.make_random_scene(nx, ny, 11, True, random.Random(1))
    cam_type, cam = rt.native.marshal_camera(sc["camera"])
    renderer.set_camera(cam_type, cam)
    lin0, img0 = renderer.render(nx, ny, 8, 50, seed=5)
    block = renderer.host_empty((2, ny, nx, 3), np.uint8)
    block[:] = 7
    plin = renderer.host_empty((ny, nx, 3), np.float32)
    plin[:] = -1.0
    renderer.render(nx, ny, 8, 50, seed=5, out_linear=plin, out_rgb8=block[1])
    assert np.all(block[0] == 7)
    assert (block[1] != img0).mean() < 1e-4
    assert np.allclose(plin, lin0, rtol=1e-5, atol=1e-6)
    _, only_img = renderer.render(nx, ny, 8, 50, seed=5, linear=False, out_rgb8=block[0])
    assert np.shares_memory(only_img, block[0]) and (block[0] != img0).mean() < 1e-4


@pytest.mark.parametrize("variant", [rt.native.RT_VARIANT_WAVEFRONT, rt.native.RT_VARIANT_MEGAKERNEL])
def test_trace_paths_against_the_plain_python_loop(renderer, random_scene_flat, variant):
    """The PRODUCTION kernels against the second, independent restatement (tests/test_second_restatement.py: `pixel` + `color` as a
    plain Python loop over numpy float64 pieces transcribed from the Clojure, fed the same Philox uniforms) — no oracle in between.
    FP32 shading against float64: the same bounce count and termination for nearly every path, radiance to 1e-3."""
    from test_second_restatement import np_color_of_sample

    flat, cam_type, cam = random_scene_flat
    nx, ny = 1200, 800
    renderer.set_scene(flat)
    renderer.set_camera(cam_type, cam)
    rng = np.random.default_rng(31)
    n = 400
    pix = rng.integers(0, nx * ny, n).astype(np.int32)
    pix[:150] = (rng.integers(250, 550, 150) * nx + rng.integers(200, 1000, 150)).astype(np.int32)   # the spheres in the middle of the frame
    smp = rng.integers(0, 4096, n).astype(np.int32)
    g_rad, g_nr, g_term, _ = renderer.trace_paths(nx, ny, pix, smp, 50, seed=123, variant=variant)
    same = close = 0
    for q in range(n):
        w_rad, w_nr, w_term = np_color_of_sample(flat, cam_type, cam, nx, ny, int(pix[q]), int(smp[q]), 123, 50)
        if (g_nr[q], g_term[q]) == (w_nr, w_term):
            same += 1
            close += bool(np.abs(g_rad[q] - w_rad).max() <= 1e-3 * max(np.abs(w_rad).max(), 1e-2))
    assert same >= 0.985 * n, f"only {same} of {n} paths take the same bounces to the same end"
    assert close >= 0.99 * same, f"radiance differs on {same - close} of {same} paths"


def test_set_scene_descriptor_cache_sees_in_place_changes(random_scene_flat):
    """native.Renderer keeps the ctypes descriptors of an unchanged FlatScene object between rt_set_scene calls; they point at the
    caller's own arrays, so a value changed IN PLACE (and an array replaced by a new one) must reach the next upload."""
    import copy

    flat0, cam_type, cam = random_scene_flat
    flat = copy.deepcopy(flat0)
    rng = np.random.default_rng(41)
    n = 4000
    o = np.tile(np.array([[13.0, 2.0, 3.0]], np.float32), (n, 1))
    d = (rng.normal(size=(n, 3)) * 0.15 + np.array([-13.0, -2.0, -3.0])).astype(np.float32)
    with rt.native.Renderer([0]) as r:
        r.set_camera(cam_type, cam)
        r.set_scene(flat)
        assert r._scene_desc is not None                               # this scene needs no converted copies: descriptors kept
        t0, id0 = r.trace_primary(o, d, None, 0.001)
        k = int(np.bincount(id0[id0 >= 0]).argmax())                   # the sphere most of these rays hit (the big one at (4, 1, 0))
        r.set_scene(flat)
        t1, id1 = r.trace_primary(o, d, None, 0.001)
        assert np.array_equal(t0, t1) and np.array_equal(id0, id1)
        flat.center0_r[k, 1] += 0.75                                   # in place: same array object, same descriptors
        r.set_scene(flat)
        t2, id2 = r.trace_primary(o, d, None, 0.001)
        flat.center0_r = flat.center0_r.copy()                         # a NEW array object: the descriptors are rebuilt
        flat.center0_r[k, 1] -= 0.75
        r.set_scene(flat)
        t3, id3 = r.trace_primary(o, d, None, 0.001)
    with rt.native.Renderer([0]) as fresh:
        moved = copy.deepcopy(flat0)
        moved.center0_r[k, 1] += 0.75
        fresh.set_camera(cam_type, cam)
        fresh.set_scene(moved)
        t_ref, id_ref = fresh.trace_primary(o, d, None, 0.001)
    assert not np.array_equal(id0, id2)                                # the move changed which rays hit what
    assert np.array_equal(t2, t_ref) and np.array_equal(id2, id_ref)
    assert np.array_equal(t3, t0) and np.array_equal(id3, id0)
