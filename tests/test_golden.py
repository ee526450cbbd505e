"""The committed fixtures of tests/golden/ (see make_golden.py for what they are and how they were made).

CPU (`-m "not gpu"`): the oracle reproduces them bit for bit and satisfies the reference's own known answers.
GPU (`-m gpu`): the CUDA path, through the C ABI, hits the same numbers — closest-hit t bit-identical, ids equal,
shading within 2e-4, a converged render within the Monte-Carlo noise the two fixture renders measure.
"""
import json
import math
import os
import random

import numpy as np
import pytest

import oracle
import raytrace_clj_b200 as rt

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FMAX = float(np.finfo(np.float32).max)


@pytest.fixture(scope="module")
def fixtures():
    return {k: np.load(os.path.join(GOLDEN, k + ".npz")) for k in ("hits_random_scene", "shade_random_scene", "image_random_scene")}


def test_oracle_reproduces_golden_hits_and_shading(random_scene_flat, fixtures):
    flat, cam_type, cam = random_scene_flat
    S = oracle.Scene(flat)
    h = fixtures["hits_random_scene"]
    t, ids = S.hit(h["origins"], h["dirs"], h["times"], 0.001, FMAX)
    assert np.array_equal(ids, h["ids"]) and np.array_equal(t, h["t"])
    assert 0.9 < (ids >= 0).mean() < 1.0 and len(np.unique(ids)) > 200      # the vectors exercise most of the scene
    s = fixtures["shade_random_scene"]
    ref = S.shade_batch(s["origins"], s["dirs"], s["times"], s["hit_id"], s["ball"], s["u01"])
    assert np.array_equal(ref["flags"], s["out_flags"])
    for k in ("origin", "dir", "atten", "emitted"):
        assert np.array_equal(ref[k].astype(np.float32), s["out_" + k]), k


def test_oracle_reproduces_golden_image(random_scene_flat, fixtures):
    flat, cam_type, _ = random_scene_flat
    g = fixtures["image_random_scene"]
    nx, ny, ns = int(g["nx"]), int(g["ny"]), 64          # a 64-sample prefix is enough to pin the sample stream...
    sc = rt.scene.make_random_scene(nx, ny, 11, True, random.Random(1))
    _, cam = rt.native.marshal_camera(sc["camera"])
    S = oracle.Scene(flat)
    part = S.render_accumulate(cam_type, cam, nx, ny, 0, ns, int(g["depth"]), seed=11)[0] / ns
    full = g["render_a"].astype(np.float64)
    # ...statistically: the prefix estimates the same image (noise sqrt(512/64) times the fixture's own)
    noise = math.sqrt(((g["render_a"].astype(np.float64) - g["render_b"]) ** 2).mean() / 2)
    assert math.sqrt(((part - full) ** 2).mean()) < 1.3 * noise * math.sqrt(int(g["spp"]) / ns + 1)


def test_reference_known_answers_hold_for_the_oracle():
    k = json.load(open(os.path.join(GOLDEN, "kat_reference_tests.json")))
    r = k["sphere_radius"]
    for c in k["grid_centres"]:
        c = np.array(c)
        for d in k["directions"]:
            d = np.array(d)
            hit = oracle.sphere_hit(c, r, c + d, -d, 0.0, FMAX)
            assert hit is not None and hit["t"] == pytest.approx(1.0 - r / np.linalg.norm(d), rel=1e-12)
            assert oracle.sphere_hit(c, r, c + d, d, 0.0, FMAX) is None
        for case in k["per_centre_hits"]:
            assert oracle.sphere_hit(c, r, c + np.array(case["origin_offset"], float), case["direction"], 0.0, FMAX) is not None, case["name"]
    pp = k["point_at_parameter"]
    for case in pp["cases"]:
        assert np.array_equal(oracle.point_at_parameter(pp["origin"], pp["direction"], case["t"]), np.array(case["p"], float))
    ct = k["center_at_time"]
    for case in ct["cases"]:
        assert np.array_equal(oracle.center_at_time(ct["c0"], ct["t0"], ct["c1"], ct["t1"], case["t"]), np.array(case["c"], float))


@pytest.mark.gpu
def test_gpu_matches_golden(random_scene_flat, fixtures):
    flat, cam_type, cam = random_scene_flat
    with rt.native.Renderer([0]) as r:
        r.set_scene(flat)
        h = fixtures["hits_random_scene"]
        t, ids = r.trace_primary(h["origins"], h["dirs"], h["times"], 0.001, FMAX)
        assert np.array_equal(ids, h["ids"])
        hit = h["ids"] >= 0
        assert np.all(np.abs(t[hit] - h["t"][hit]) <= 1e-5 * np.abs(h["t"][hit]))      # north_star tolerance
        assert np.array_equal(t[hit], h["t"][hit]) and np.all(np.isinf(t[~hit]))        # in fact bit-identical
        s = fixtures["shade_random_scene"]
        g = r.shade_batch(s["origins"], s["dirs"], s["times"], s["hit_id"], s["hit_t"], s["ball"], s["u01"])
        same = g["flags"] == s["out_flags"]
        assert (~same).mean() < 1e-3
        assert np.allclose(g["emitted"][same], s["out_emitted"][same], atol=2e-4)
        cont = same & (s["out_flags"] == 1)
        close = np.all(np.abs(g["dir"][cont] - s["out_dir"][cont]) <= 2e-3 * np.maximum(1.0, np.linalg.norm(s["out_dir"][cont], axis=1))[:, None], axis=1)
        assert close.mean() > 0.995            # Schlick / TIR threshold cases may take the other branch in FP32
        idx = np.nonzero(cont)[0][close]
        assert np.allclose(g["origin"][idx], s["out_origin"][idx], rtol=1e-6, atol=2e-4)
        att_ok = np.all(np.abs(g["atten"][idx] - s["out_atten"][idx]) <= 1e-6, axis=1)
        assert att_ok.mean() > 0.995           # a checker sample on a sine zero crossing may flip colour in FP32
        # converged image: RMSE(gpu, fixture) within the noise the two fixture renders measure
        im = fixtures["image_random_scene"]
        nx, ny, ns, depth = int(im["nx"]), int(im["ny"]), int(im["spp"]), int(im["depth"])
        sc = rt.scene.make_random_scene(nx, ny, 11, True, random.Random(1))
        cam_type, cam_small = rt.native.marshal_camera(sc["camera"])
        r.set_camera(cam_type, cam_small)
        a, b = im["render_a"].astype(np.float64), im["render_b"].astype(np.float64)
        r_oo = math.sqrt(((a - b) ** 2).mean())
        for variant in (0, 1):
            lin, _ = r.render(nx, ny, ns, depth, seed=5 + variant, variant=variant)
            cross = max(math.sqrt(((lin - a) ** 2).mean()), math.sqrt(((lin - b) ** 2).mean()))
            assert cross <= 1.25 * r_oo, (variant, cross, r_oo)
            mean_diff = np.abs(lin.mean(axis=(0, 1)) - (a + b).mean(axis=(0, 1)) / 2)
            sigma = np.sqrt(((a - b) ** 2).mean(axis=(0, 1)) / 2)
            assert np.all(mean_diff <= 3 * sigma / math.sqrt(nx * ny) * math.sqrt(1.5) + 2e-4), (variant, mean_diff)
