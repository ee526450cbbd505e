import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def random_scene_flat():
    """make-random-scene (scene.clj:318-412), n=11, moving=true, scene seed 1, marshalled."""
    import random

    import raytrace_clj_b200 as rt

    sc = rt.scene.make_random_scene(1200, 800, 11, True, random.Random(1))
    flat = rt.native.marshal_world(sc["world"])
    cam_type, cam = rt.native.marshal_camera(sc["camera"])
    return flat, cam_type, cam
