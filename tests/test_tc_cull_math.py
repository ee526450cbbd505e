"""CPU check of the tensor-core cull's arithmetic (raytrace_clj_b200/csrc/rt_cull_tc.cuh): the bilinear form, the hi/lo
TF32 split of its 32 K-slots and the rounding budget, restated in numpy.  No GPU: products of TF32 parts are exact in
double, so what is checked here is everything EXCEPT the tensor core's own accumulation error — which the budget covers
with 16 u sum|terms| (measured <= 7 u by csrc/tcprobe.cu) and which this test subtracts explicitly:

    for every (ray, sphere) whose LINE really meets the sphere:   sum_k ray_k * sphere_k  -  16 u sum_k |ray_k * sphere_k|  >=  0
"""
import numpy as np

U = 2.0 ** -24
CULL_EPS = np.float32(2.0 ** -17)
RAY_DEFLATE_TC = np.float32(1.0) - np.float32(5.0e-6)


def tf32_rn(x):
    b = np.asarray(x, np.float32).view(np.uint32)
    return ((b + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)


def split2(x):
    x = np.asarray(x, np.float32)
    hi = tf32_rn(x)
    lo = tf32_rn((x - hi).astype(np.float32))
    return hi, lo


def f32(x):
    return np.asarray(x, np.float64).astype(np.float32)


def fma32(a, b, c):
    return f32(a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64))


def ray_slots(o, d):
    """tc_produce_ray: [n, 32] float32."""
    o = o.astype(np.float32); d = d.astype(np.float32)
    aa = fma32(d[:, 2], d[:, 2], fma32(d[:, 1], d[:, 1], f32(d[:, 0].astype(np.float64) * d[:, 0])))
    inv = f32(1.0 / np.sqrt(aa.astype(np.float64)))                     # rsqrtf (2 ulp on the device; covered by eps)
    qq = fma32(o[:, 2], o[:, 2], fma32(o[:, 1], o[:, 1], f32(o[:, 0].astype(np.float64) * o[:, 0])))
    s = f32(inv.astype(np.float64) * np.float64(np.float32(1.0) + np.float32(0.5) * CULL_EPS))
    h = f32(d.astype(np.float64) * s[:, None])
    P = fma32(o[:, 2], h[:, 2], fma32(o[:, 1], h[:, 1], f32(o[:, 0].astype(np.float64) * h[:, 0])))
    g0 = fma32(P, P, -f32(qq.astype(np.float64) * RAY_DEFLATE_TC))
    g = f32(2.0 * fma32(-P[:, None] * np.ones(3, np.float32), h, o).astype(np.float64))
    n = len(o)
    f = np.zeros((n, 32), np.float32)
    hi, lo = split2(g0); f[:, 0] = hi; f[:, 1] = lo
    f[:, 2] = f[:, 3] = f[:, 7] = 1.0
    for axis, base in ((0, 4), (1, 8), (2, 12)):
        hi, lo = split2(g[:, axis]); f[:, base] = hi; f[:, base + 1] = hi; f[:, base + 2] = lo
    q = lambda a, b: f32(h[:, a].astype(np.float64) * h[:, b])
    for (a, b), base in (((0, 0), 16), ((1, 1), 19), ((2, 2), 22), ((0, 1), 25), ((0, 2), 28)):
        hi, lo = split2(q(a, b)); f[:, base] = hi; f[:, base + 1] = hi; f[:, base + 2] = lo
    hi, lo = split2(q(1, 2)); f[:, 31] = hi; f[:, 11] = hi; f[:, 15] = lo
    return f


def sphere_slots(c, r):
    """build_cull_records + tc::sphere_slots for static plain spheres: [m, 32] float32.  c = centre as stored (float32)."""
    c = c.astype(np.float32).astype(np.float64)
    cc = (c * c).sum(axis=1)
    eps = 2.0 ** -20
    W = r * r * (1.0 + eps) + 96.0 * U * cc + 24.0 * U * r * r - cc
    sp2 = lambda x: (lambda hi: (hi, tf32_rn(f32(x - hi.astype(np.float64)))))(tf32_rn(f32(x)))
    wh = tf32_rn(f32(W)); wl = tf32_rn(f32(W - wh.astype(np.float64))); wl2 = tf32_rn(f32(W - wh.astype(np.float64) - wl.astype(np.float64)))
    m = len(c)
    s = np.zeros((m, 32), np.float32)
    s[:, 0] = s[:, 1] = 1.0
    s[:, 2] = wh; s[:, 3] = wl; s[:, 7] = wl2
    for axis, base in ((0, 4), (1, 8), (2, 12)):
        hi, lo = sp2(c[:, axis]); s[:, base] = hi; s[:, base + 1] = lo; s[:, base + 2] = hi
    for (a, b), base in (((0, 0), 16), ((1, 1), 19), ((2, 2), 22)):
        hi, lo = sp2(c[:, a] * c[:, b]); s[:, base] = hi; s[:, base + 1] = lo; s[:, base + 2] = hi
    for (a, b), base in (((0, 1), 25), ((0, 2), 28)):
        hi, lo = sp2(2.0 * c[:, a] * c[:, b]); s[:, base] = hi; s[:, base + 1] = lo; s[:, base + 2] = hi
    hi, lo = sp2(2.0 * c[:, 1] * c[:, 2]); s[:, 31] = hi; s[:, 11] = lo; s[:, 15] = hi
    return s


def exact_disc(o, d, c, r):
    """(u.(o-c))^2 - |o-c|^2 + r^2 with u = d / |d|, in double, on the float32 inputs."""
    o = o.astype(np.float32).astype(np.float64); d = d.astype(np.float32).astype(np.float64)
    c = c.astype(np.float32).astype(np.float64)
    u = d / np.linalg.norm(d, axis=1)[:, None]
    oc = o - c
    b = (u * oc).sum(axis=1)
    return b * b - (oc * oc).sum(axis=1) + r * r


def tangent_pairs(rng, n, origin_scale, centre_scale, r_lo, r_hi, jitter):
    c = rng.normal(size=(n, 3)) * centre_scale
    r = np.exp(rng.uniform(np.log(r_lo), np.log(r_hi), n))
    o = rng.normal(size=(n, 3)) * origin_scale
    to_c = c - o
    perp = np.cross(to_c, rng.normal(size=(n, 3)))
    perp /= np.linalg.norm(perp, axis=1)[:, None]
    target = c + perp * (r * (1.0 + rng.normal(scale=jitter, size=n)))[:, None]
    d = (target - o) * rng.uniform(0.1, 4.0, size=(n, 1))
    return o, d, c, r


def test_every_slot_is_tf32_representable():
    rng = np.random.default_rng(0)
    o, d, c, r = tangent_pairs(rng, 2000, 10.0, 10.0, 0.05, 5.0, 1e-3)
    for m in (ray_slots(o, d), sphere_slots(c, r)):
        assert np.array_equal(m, tf32_rn(m)) or np.array_equal(m.view(np.uint32) & 0x1FFF, np.zeros_like(m, np.uint32))


def test_bilinear_form_is_the_line_sphere_discriminant():
    """Away from tangency the 32-slot product reproduces the discriminant to the budget's accuracy."""
    rng = np.random.default_rng(1)
    o = rng.normal(size=(4000, 3)) * 8.0
    d = rng.normal(size=(4000, 3))
    c = rng.normal(size=(4000, 3)) * 8.0
    r = rng.uniform(0.1, 3.0, 4000)
    disc = (ray_slots(o, d).astype(np.float64) * sphere_slots(c, r)).sum(axis=1)
    ref = exact_disc(o, d, c, r)
    oo = (o.astype(np.float32).astype(np.float64) ** 2).sum(axis=1); cc = (c.astype(np.float32).astype(np.float64) ** 2).sum(axis=1)
    # the form is deliberately biased upwards: eps on the direction, the deflated |o|^2, the inflated radius
    slack = 2.0 ** -15 * (oo + cc + r * r)          # eps = 2^-17 on (u.oc)^2 <= 2 (|o|^2 + |c|^2) dominates
    assert np.all(disc >= ref - 1e-12)
    assert np.all(disc <= ref + slack)


def test_never_under_reports_even_after_the_accumulation_budget():
    rng = np.random.default_rng(2)
    worst = 0.0
    total = 0
    for origin_scale, centre_scale, r_lo, r_hi in ((1, 1, 0.05, 1), (15, 15, 0.2, 1.0), (100, 10, 0.2, 5), (1000, 1000, 0.1, 50),
                                                   (3000, 15, 0.2, 1000), (10, 1000, 100, 1000)):
        for jitter in (0.0, 3e-7, 1e-5, 1e-3):
            o, d, c, r = tangent_pairs(rng, 20_000, origin_scale, centre_scale, r_lo, r_hi, jitter)
            R = ray_slots(o, d).astype(np.float64)
            S = sphere_slots(c, r).astype(np.float64)
            terms = R * S
            disc = terms.sum(axis=1)
            budget = 16.0 * U * np.abs(terms).sum(axis=1)            # what the tensor core's accumulation may lose
            hit = exact_disc(o, d, c, r) >= 0.0
            total += int(hit.sum())
            bad = hit & (disc - budget < 0.0)
            assert not bad.any(), (origin_scale, centre_scale, jitter, int(bad.sum()))
            worst = max(worst, float(np.max(np.where(hit, budget / np.maximum(disc, 1e-300), 0.0))))
    assert total > 100_000


def test_selectivity_stays_close_to_the_exact_test():
    """The inflation is a few ulps of the magnitudes involved: on scene-scale data only a sliver of extra pairs passes."""
    rng = np.random.default_rng(3)
    o = rng.normal(size=(3000, 3)) * 10.0
    d = rng.normal(size=(3000, 3))
    c = rng.uniform(-11, 11, size=(400, 3)); c[:, 1] = 0.2
    r = np.full(400, 0.2)
    R = ray_slots(o, d).astype(np.float64)
    S = sphere_slots(c, r).astype(np.float64)
    disc = R @ S.T
    oo = np.repeat(o, 400, axis=0); dd = np.repeat(d, 400, axis=0); ccs = np.tile(c, (3000, 1)); rr = np.tile(r, 3000)
    ref = exact_disc(oo, dd, ccs, rr).reshape(3000, 400)
    assert not ((ref >= 0) & (disc < 0)).any()
    assert (disc >= 0).sum() <= 1.10 * (ref >= 0).sum() + 5      # r = 0.2 at |o| ~ 17: the budget is ~10 % of r^2


def test_dead_rays_and_padding_rows_never_survive():
    """A dead ray (tile slot beyond the queue) and a padding row (feature row beyond the leaf list) must give a NEGATIVE
    discriminant against anything — in particular against each other (0 x 0 = +0 reads as "candidate": the bug that made
    the last, partial ray tile of a launch flood the candidate lists on scenes with few leaves)."""
    DEAD = np.float32(-1.0e30)
    dead = np.zeros(32, np.float32); dead[0] = DEAD; dead[2] = 1.0           # tc_produce_ray, dead branch
    pad = np.zeros(32, np.float32); pad[0] = 1.0; pad[2] = DEAD              # tc::padding_slots
    rng = np.random.default_rng(4)
    o, d, c, r = tangent_pairs(rng, 500, 50.0, 50.0, 0.1, 30.0, 1e-3)
    R = ray_slots(o, d).astype(np.float64); S = sphere_slots(c, r).astype(np.float64)
    assert float(dead.astype(np.float64) @ pad) < -1e29
    assert np.all(R @ pad.astype(np.float64) < -1e29)
    assert np.all(S @ dead.astype(np.float64) < -1e29)
