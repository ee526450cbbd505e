"""Image writers for the front end.  The reference saves through imagez `save`
(core.clj:112), which picks a javax.imageio writer by extension and has none for `.ppm`; the
documented CLI (`lein run out.ppm …`) therefore needs a PPM writer."""
from __future__ import annotations

import numpy as np


def write_ppm(path, rgb8: np.ndarray) -> None:
    """Binary P6; rgb8 is [ny, nx, 3] uint8, row 0 = top."""
    ny, nx, _ = rgb8.shape
    with open(path, "wb") as f:
        f.write(f"P6\n{nx} {ny}\n255\n".encode("ascii"))
        f.write(np.ascontiguousarray(rgb8, np.uint8).tobytes())


def read_ppm(path) -> np.ndarray:
    with open(path, "rb") as f:
        data = f.read()
    parts = data.split(b"\n", 3)
    assert parts[0] == b"P6"
    nx, ny = (int(x) for x in parts[1].split())
    return np.frombuffer(parts[3], np.uint8).reshape(ny, nx, 3)


def save(path, rgb8: np.ndarray) -> None:
    """Pick the writer by extension like imagez does: .ppm here, anything else through PIL."""
    if str(path).lower().endswith(".ppm"):
        write_ppm(path, rgb8)
    else:
        from PIL import Image  # noqa: WPS433 (optional dependency, only for non-ppm output)
        Image.fromarray(rgb8, "RGB").save(path)


def load(path) -> np.ndarray:
    """uint8 [ny, nx, 3], row 0 = top: .ppm here, anything else through PIL (texture.clj:135-138 load-image)."""
    if str(path).lower().endswith(".ppm"):
        return read_ppm(path)
    from PIL import Image  # noqa: WPS433
    return np.asarray(Image.open(path).convert("RGB"), np.uint8)
