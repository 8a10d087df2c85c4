"""A data-generation script in the style of the reference's scripts/datagen_denoise.py and datagen_colmap.py, written for
the tests: it drives the embedded `volpy` module through the same calls in the same order (Renderer(), init(), draw(),
Volume(path), commit(), member assignment, Environment(path).strength, AABB("density") vec3 arithmetic, static camera
properties, render(spp), fbo_data() -> numpy flip/transpose, draw(), save_with_alpha(), colmap_* helpers, shutdown()).
Run as:  volren tests/scripts/datagen_like.py -w 48 -h 48 --render      (outputs go to $VOLREN_TEST_OUT)
"""
import math
import os
import random

import numpy as np

import volpy

if __name__ == "__main__":
    ROOT_DIR = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    ASSETS = os.path.join(ROOT_DIR, "tests", "golden", "assets")
    OUT = os.environ.get("VOLREN_TEST_OUT", ".")
    N_IMAGES = 2
    N_SAMPLES_TARGET = 32

    renderer = volpy.Renderer()
    renderer.init()
    renderer.draw()
    random.seed(42)

    SIZE = renderer.resolution()
    inputs = np.zeros((N_IMAGES, 3, SIZE.y, SIZE.x), np.float16)
    targets = np.zeros((N_IMAGES, 3, SIZE.y, SIZE.x), np.float16)

    def uniform_sample_sphere():
        z = 1.0 - 2.0 * random.random()
        r = math.sqrt(max(0.0, 1.0 - z * z))
        phi = 2.0 * math.pi * random.random()
        return volpy.vec3(r * math.cos(phi), r * math.sin(phi), z)

    qvecs, tvecs = [], []
    for i in range(N_IMAGES):
        renderer.volume = volpy.Volume(os.path.join(ASSETS, "smoke.brick"))
        renderer.commit()
        renderer.albedo = volpy.vec3(random.random(), random.random(), random.random())
        renderer.phase = -0.9 + random.random() * 1.8
        renderer.density_scale = 0.5 + random.random() * 5
        renderer.environment = volpy.Environment(os.path.join(ASSETS, "table_mountain_2_puresky_1k.hdr"))
        renderer.environment.strength = 0.5 + random.random() * 10
        renderer.show_environment = i == 0
        renderer.transferfunc = None
        bb_min, bb_max = renderer.volume.AABB("density")
        center = bb_min + (bb_max - bb_min) * 0.5
        radius = (bb_max - center).length()
        renderer.cam_pos = center + uniform_sample_sphere() * radius
        renderer.cam_dir = (center + uniform_sample_sphere() * radius * 0.1 - renderer.cam_pos).normalize()
        renderer.cam_fov = 25 + random.random() * 70
        renderer.seed = random.randint(0, 2**31)
        renderer.bounces = random.randint(1, 16)
        renderer.render(random.randint(1, 4))
        data = np.flip(np.array(renderer.fbo_data()), axis=0)
        inputs[i] = np.transpose(data.astype(np.float16), [2, 1, 0])
        renderer.draw()
        renderer.seed = random.randint(0, 2**31)
        renderer.render(N_SAMPLES_TARGET)
        data = np.flip(np.array(renderer.fbo_data()), axis=0)
        targets[i] = np.transpose(data.astype(np.float16), [2, 1, 0])
        renderer.tonemapping = True
        renderer.draw()
        renderer.save_with_alpha(os.path.join(OUT, f"view_{i:06}.png"))
        qvecs.append(np.array(renderer.colmap_view_rot())[[3, 0, 1, 2]])
        tvecs.append(np.array(renderer.colmap_view_trans()))

    focal = renderer.colmap_focal_length()
    np.savez(os.path.join(OUT, "dataset.npz"), inputs=inputs, targets=targets, qvecs=np.array(qvecs), tvecs=np.array(tvecs), focal=focal,
             cx=renderer.resolution().x // 2, aabb=np.array([np.array(v) for v in renderer.volume.AABB("density")]))
    print("datagen_like done")
    renderer.shutdown()
