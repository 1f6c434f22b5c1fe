"""The benchmark configurations of BASELINE.json / SURVEY 8(d): cameras, lights and flags as plain data."""
from fluctus_b200 import look_at, make_params

CEILING_LIGHT_CONFERENCE = dict(pos=(0.0, 0.235, 0.0), N=(0.0, -1.0, 0.0), right=(1.0, 0.0, 0.0), up=(0.0, 0.0, 1.0), size=(0.25, 0.25), E=(200.0, 200.0, 200.0))


def conference_params(scene, width, height, max_bounces=8):
    cam = look_at((-0.80, 0.05, 0.50), (0.60, -0.08, -0.30), fov=60.0)
    return make_params(width, height, cam, scene.world_radius, len(scene.tris), light=CEILING_LIGHT_CONFERENCE, max_bounces=max_bounces)
