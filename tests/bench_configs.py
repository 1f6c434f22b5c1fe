"""The benchmark configurations of BASELINE.json / SURVEY 8(d): cameras, lights and flags as plain data."""
from fluctus_b200 import look_at, make_params

CEILING_LIGHT_CONFERENCE = dict(pos=(0.0, 0.235, 0.0), N=(0.0, -1.0, 0.0), right=(1.0, 0.0, 0.0), up=(0.0, 0.0, 1.0), size=(0.25, 0.25), E=(200.0, 200.0, 200.0))


def conference_params(scene, width, height, max_bounces=8):
    cam = look_at((-0.80, 0.05, 0.50), (0.60, -0.08, -0.30), fov=60.0)
    return make_params(width, height, cam, scene.world_radius, len(scene.tris), light=CEILING_LIGHT_CONFERENCE, max_bounces=max_bounces)


def luxball_params(scene, width, height, max_bounces=16):
    cam = look_at((0.0, 1.6, 3.3), (0.0, 0.9, 0.0), fov=55.0)
    light = dict(pos=(0.0, 3.45, 0.0), N=(0.0, -1.0, 0.0), right=(1.0, 0.0, 0.0), up=(0.0, 0.0, 1.0), size=(1.0, 1.0), E=(200.0, 200.0, 200.0))
    return make_params(width, height, cam, scene.world_radius, len(scene.tris), light=light, max_bounces=max_bounces, separate_queues=True)


def kitchen_params(scene, width, height, max_bounces=8):
    cam = look_at((2.6, 1.7, 4.4), (0.0, 1.0, 0.0), fov=60.0)
    return make_params(width, height, cam, scene.world_radius, len(scene.tris), light=False, max_bounces=max_bounces, use_env_map=True,
                       env_map_strength=30.0, separate_queues=True)


def teapot_params(scene, width, height, max_bounces=2):
    cam = dict(pos=(0, 1, 3.5), dir=(0, 0, -1), right=(1, 0, 0), up=(0, 1, 0), fov=60.0)  # Tracer::initCamera, tracer.cpp:760-776
    return make_params(width, height, cam, scene.world_radius, len(scene.tris), max_bounces=max_bounces)


CONFIGS = {"conference": conference_params, "luxball": luxball_params, "country_kitchen": kitchen_params, "teapot": teapot_params}
ENV_MAPS = {"country_kitchen": "night"}


def params_for(name, scene, width, height, max_bounces=None):
    fn = CONFIGS[name]
    return fn(scene, width, height) if max_bounces is None else fn(scene, width, height, max_bounces)
