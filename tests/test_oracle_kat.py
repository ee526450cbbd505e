"""Pins the CPU oracle against every known-answer case the reference's own tests hold for the
hot path (test/raytrace_clj/hitable_test.clj, util_test.clj) and against closed-form values
derived from the reference formulas.  CPU only."""
import itertools
import math

import numpy as np
import pytest

import oracle

FMAX = float(np.finfo(np.float32).max)

# hitable_test.clj:8-11 — 27 centres 25*{-1,0,1}^3; :13-19 — 26 directions 5*{-1,0,1}^3 \ 0
GRIDPOINTS = [25.0 * np.array(p, float) for p in itertools.product((-1, 0, 1), repeat=3)]
DIRECTIONS = [5.0 * np.array(p, float) for p in itertools.product((-1, 0, 1), repeat=3) if any(p)]


def test_point_at_parameter_util_test_44_49():
    o, d = [1, 2, 3], [4, 5, 6]
    assert np.array_equal(oracle.point_at_parameter(o, d, 0), [1, 2, 3])
    assert np.array_equal(oracle.point_at_parameter(o, d, 1), [5, 7, 9])
    assert np.array_equal(oracle.point_at_parameter(o, d, -1), [-3, -3, -3])


def test_center_at_time_hitable_test_97_103():
    pa, pb = [0, 0, 0], [1, 2, 3]
    assert np.array_equal(oracle.center_at_time(pa, 0, pb, 1, 0), pa)
    assert np.array_equal(oracle.center_at_time(pa, 0, pb, 1, 1), pb)
    assert np.array_equal(oracle.center_at_time(pa, 0, pb, 1, 0.5), [0.5, 1.0, 1.5])


@pytest.mark.parametrize("origin", GRIDPOINTS, ids=lambda p: "c%+d%+d%+d" % tuple(p / 25))
def test_sphere_hit_miss_hitable_test_23_47(origin):
    r = 1.0
    for d in DIRECTIONS:
        nz = int(np.count_nonzero(d))
        h = oracle.sphere_hit(origin, r, origin + d, -d, 0.0, FMAX)           # :32-34 "intersect ray"
        assert h is not None
        # closed form: t = 1 - 1/|dir|  (SURVEY §4): 0.8, 0.8585786437626906, 0.8845299461620749
        assert h["t"] == pytest.approx(1.0 - 1.0 / math.sqrt(25.0 * nz), rel=1e-14)
        assert np.allclose(h["normal"], d / np.linalg.norm(d), atol=1e-14)
        assert np.allclose(h["p"], origin + d / np.linalg.norm(d), atol=1e-12)
        assert oracle.sphere_hit(origin, r, origin + d, d, 0.0, FMAX) is None  # :35-36 "non-intersecting ray"
    # :37-45 grazing rays: discriminant exactly 0, t = 1
    for off, dr in (([r, r, 0], [-1, 0, 0]), ([r, r, 0], [0, -1, 0]), ([r, 0, r], [0, 0, -1])):
        h = oracle.sphere_hit(origin, r, origin + np.array(off, float), dr, 0.0, FMAX)
        assert h is not None and h["t"] == 1.0
    # :46-47 ray from the centre: near root negative -> far root 1/sqrt(3)
    h = oracle.sphere_hit(origin, r, origin, [1, 1, 1], 0.0, FMAX)
    assert h is not None and h["t"] == pytest.approx(1 / math.sqrt(3), rel=1e-14)
    assert np.allclose(h["normal"], np.ones(3) / math.sqrt(3), atol=1e-14)


@pytest.mark.parametrize("origin", GRIDPOINTS[::3], ids=lambda p: "c%+d%+d%+d" % tuple(p / 25))
def test_moving_sphere_hit_miss_hitable_test_61_83(origin):
    dest = origin + np.array([10.0, 20.0, 30.0])
    t0, t1, r = 0.1, 0.9, 1.0
    for d in DIRECTIONS:
        assert oracle.sphere_hit(origin, r, origin + d, -d, 0.0, FMAX, time=t0, center1=dest, t0=t0, t1=t1) is not None
        assert oracle.sphere_hit(origin, r, origin + d, d, 0.0, FMAX, time=t0, center1=dest, t0=t0, t1=t1) is None
    assert oracle.sphere_hit(origin, r, origin, [1, 1, 1], 0.0, FMAX, time=t0, center1=dest, t0=t0, t1=t1) is not None
    # at time t1 the sphere sits at `dest`
    h = oracle.sphere_hit(origin, r, dest + [5.0, 0, 0], [-5.0, 0, 0], 0.0, FMAX, time=t1, center1=dest, t0=t0, t1=t1)
    assert h is not None and h["t"] == pytest.approx(0.8, rel=1e-12)


def test_strict_range_and_root_order():
    # hitable.clj:196,204 strict inequalities: t == t_min / t == t_max are rejected
    assert oracle.sphere_hit([0, 0, 0], 1.0, [5, 0, 0], [-5, 0, 0], 0.8, FMAX)["t"] == pytest.approx(1.2)  # far root
    assert oracle.sphere_hit([0, 0, 0], 1.0, [5, 0, 0], [-5, 0, 0], 0.0, 0.8) is None
    assert oracle.sphere_hit([0, 0, 0], 1.0, [5, 0, 0], [-5, 0, 0], 0.0, 0.8000001)["t"] == pytest.approx(0.8)


def test_get_sphere_uv_hitable_128_139():
    assert np.allclose(oracle.get_sphere_uv([1, 0, 0]), [0.5, 0.5])
    assert np.allclose(oracle.get_sphere_uv([0, 1, 0]), [0.5, 1.0])
    assert np.allclose(oracle.get_sphere_uv([0, -1, 0]), [0.5, 0.0])
    assert np.allclose(oracle.get_sphere_uv([0, 0, 1]), [0.25, 0.5])
    assert np.allclose(oracle.get_sphere_uv([-1, 0, 0]), [0.0, 0.5])  # atan2(0,-1) = +pi


def test_reflect_refract_schlick_shader_6_20_69_74():
    assert np.allclose(oracle.reflect([1, -1, 0], [0, 1, 0]), [1, 1, 0])
    # un-normalised input is kept un-normalised by reflect (Dielectric passes raw directions)
    assert np.allclose(oracle.reflect([2, -2, 0], [0, 1, 0]), [2, 2, 0])
    # straight through along -n: direction unchanged (unit length)
    assert np.allclose(oracle.refract([0, -3, 0], [0, 1, 0], 1 / 1.5), [0, -1, 0])
    # Snell: sin(t) = sin(i)/1.5
    v = np.array([math.sin(0.5), -math.cos(0.5), 0])
    out = oracle.refract(v, [0, 1, 0], 1 / 1.5)
    assert out[0] == pytest.approx(math.sin(0.5) / 1.5, rel=1e-14)
    assert np.linalg.norm(out) == pytest.approx(1.0, rel=1e-14)
    # total internal reflection: glass -> air beyond the critical angle
    v = np.array([math.sin(1.0), math.cos(1.0), 0])
    assert oracle.refract(v, [0, -1, 0], 1.5) is None
    r0 = ((1 - 1.5) / (1 + 1.5)) ** 2
    assert oracle.schlick(1.0, 1.5) == pytest.approx(r0)
    assert oracle.schlick(0.0, 1.5) == pytest.approx(1.0)
    assert oracle.schlick(0.5, 1.5) == pytest.approx(r0 + (1 - r0) * 0.5 ** 5)


def test_camera_known_answers_survey_8a():
    """camera.clj:50-66 for make-random-scene at 1200x800 (values computed in the survey)."""
    c = oracle.thin_lens_camera([13, 2, 3], [0, 0, 0], [0, 1, 0], 20, 1200.0 / 800.0, 0.0, 10.0, 0.0, 1.0)
    assert np.allclose(c[12:15], [0.22485950669875843, 0, -0.9743911956946198], rtol=0, atol=1e-15)
    assert np.allclose(c[15:18], [-0.14445336159384606, 0.9889499370655614, -0.0333353911370414], atol=1e-15)
    assert np.allclose(c[18:21], [0.9636241116594315, 0.14824986333222023, 0.22237479499833035], atol=1e-15)
    assert np.allclose(c[3:6], [3.023737165939192, -1.2262841980681713, 3.4122032022021487], atol=1e-14)
    assert np.allclose(c[6:9], [1.1894639369936078, 0, -5.1543437269723], atol=1e-14)
    assert np.allclose(c[9:12], [-0.5094205020606202, 3.487571129491938, -0.11755857739860466], atol=1e-14)
    o, d, t = oracle.get_ray(1, c, 0.5, 0.5, disk=(0.3, -0.2), time_u=0.25)
    assert np.allclose(o, [13, 2, 3]) and t == 0.25                   # aperture 0: the disk draw is discarded
    assert np.allclose(d, -10.0 * c[18:21], atol=1e-6) and np.linalg.norm(d) == pytest.approx(10.0, rel=1e-7)
    # python host mirror builds the same record
    import raytrace_clj_b200 as rt
    cm = rt.camera.thin_lens_camera(lookfrom=rt.util.vec3(13, 2, 3), lookat=rt.util.vec3(0, 0, 0),
                                    vup=rt.util.vec3(0, 1, 0), vfov=20, aspect=1200.0 / 800.0, aperture=0.0,
                                    focus_dist=10.0, t0=0.0, t1=1.0)
    assert np.allclose(np.concatenate([cm.origin, cm.lleft, cm.horiz, cm.vert, cm.u, cm.v, cm.w]), c[:21], atol=1e-14)


def test_pinhole_camera_camera_18_33():
    c = oracle.pinhole_camera([0, 0, 0], [0, 0, -1], [0, 1, 0], 90, 2.0)
    assert np.allclose(c[3:6], [-2, -1, -1]) and np.allclose(c[6:9], [4, 0, 0]) and np.allclose(c[9:12], [0, 2, 0])
    o, d, t = oracle.get_ray(0, c, 0.5, 0.5)
    assert np.allclose(o, 0) and np.allclose(d, [0, 0, -1]) and t == 0


def test_textures_texture_14_50(random_scene_flat):
    flat, _, _ = random_scene_flat
    S = oracle.Scene(flat)
    sky_tex = int(flat.mat_tex[flat.material_id[np.nonzero(flat.sphere_flags & 1)[0][0]]])
    # UVGradient with co=cu=(1,1,1), cv=cuv=(.5,.7,1): white at v=1 (zenith), blue at v=0
    assert np.allclose(S.tex_sample(sky_tex, 0.3, 1.0, [0, 0, 0]), [1, 1, 1])
    assert np.allclose(S.tex_sample(sky_tex, 0.3, 0.0, [0, 0, 0]), [0.5, 0.7, 1.0])
    assert np.allclose(S.tex_sample(sky_tex, 0.9, 0.5, [0, 0, 0]), [0.75, 0.85, 1.0])
    chk = int(np.nonzero(flat.tex_type == 2)[0][0])
    # sin(10x) sin(10y) sin(10z) < 0 -> tex0 (.2,.3,.1) else tex1 (.9,.9,.9)
    assert np.allclose(S.tex_sample(chk, 0, 0, [0.1, 0.1, 0.1]), [0.9, 0.9, 0.9])
    assert np.allclose(S.tex_sample(chk, 0, 0, [0.1, -0.1, 0.1]), [0.2, 0.3, 0.1])


def test_hitlist_closest_and_first_wins_ties():
    """hitable.clj:15-26: shrinking t-max, strictly closer wins => first item wins exact ties."""
    import raytrace_clj_b200 as rt
    from raytrace_clj_b200.util import vec3
    m = rt.shader.lambertian(albedo=rt.texture.constant(color=vec3(.5, .5, .5)))
    world = rt.hitable.hitlist(items=[
        rt.hitable.sphere(center=vec3(0, 0, -10), radius=1, material=m),
        rt.hitable.sphere(center=vec3(0, 0, -5), radius=1, material=m),
        rt.hitable.sphere(center=vec3(0, 0, -5), radius=1, material=m),   # exact duplicate of #1
        rt.hitable.sphere(center=vec3(0, 0, 5), radius=1, material=m),    # behind the ray
    ])
    S = oracle.Scene(rt.native.marshal_world(world))
    t, ids = S.hit([[0, 0, 0]], [[0, 0, -1]], [0.0])
    assert ids[0] == 1 and t[0] == 4.0
    t, ids = S.hit([[0, 0, 0]], [[0, 1, 0]], [0.0])
    assert ids[0] == -1 and math.isinf(t[0])
    # un-normalised direction: t is in units of |d| (util.clj:13-16)
    t, ids = S.hit([[0, 0, 0]], [[0, 0, -4]], [0.0])
    assert ids[0] == 1 and t[0] == 1.0


def test_resolve_core_52_57():
    s = np.zeros((2, 3, 3))
    s[0, 0] = [4 * 1.0, 4 * 0.25, 0.0]          # bottom-left pixel (j=0)
    s[1, 2] = [4 * 7.0, float("nan"), 4 * 0.9999]
    img = oracle.resolve(s, 4)
    assert img.shape == (2, 3, 3)
    assert list(img[1, 0]) == [255, 127, 0]      # j=0 is written to row ny-1 (core.clj:105); int(.5*255.99)=127
    assert list(img[0, 2]) == [255, 0, int(math.sqrt(0.9999) * 255.99)]   # clamp at 255.99; (int NaN) = 0


def test_tiled_coords_core_59_71():
    import raytrace_clj_b200 as rt
    tiles = rt.core.tiled_coords(5, 3, 4)
    assert len(tiles) == 4 and tiles[0]["chunk"] == [(0, 0), (1, 0), (2, 0), (3, 0)]
    assert tiles[1]["chunk"][0] == (4, 0) and tiles[1]["chunk"][1] == (0, 1)     # j outer, i inner
    assert tiles[-1]["chunk"] == [(2, 2), (3, 2), (4, 2)] and tiles[-1]["pct"] == 1.0
    assert len(rt.core.tiled_coords(200, 100, 32)) == 625
