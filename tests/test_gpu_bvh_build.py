"""GPU tests of the hierarchy builder flx_build_bvh: node and index arrays bit-identical to the CPU
restatement (oracle/bvh_oracle.c) on procedural, reference-asset and degenerate inputs; and the wavefront path rendering
through the built tree stays in bit-for-bit lockstep with the oracle rendering through the same tree."""
import numpy as np
import pytest

from fluctus_b200 import CLContext, SceneData
from fluctus_b200.scene import make_room_scene, room_params

from conftest import scene_blob
from parity_util import run_lockstep, validate_bvh
from test_bvh_build_cpu import teapot_scene, tri_soup
from test_gpu_parity import oracle_ctx

pytestmark = pytest.mark.gpu


def same_tree(gpu_nodes, gpu_idx, cpu_nodes, cpu_idx, what):
    assert np.array_equal(gpu_idx, cpu_idx), "%s: index lists differ" % what
    assert len(gpu_nodes) == len(cpu_nodes), "%s: %d vs %d nodes" % (what, len(gpu_nodes), len(cpu_nodes))
    a, b = gpu_nodes.view(np.uint8).reshape(len(gpu_nodes), 48), cpu_nodes.view(np.uint8).reshape(len(cpu_nodes), 48)
    bad = np.flatnonzero((a != b).any(axis=1))
    assert len(bad) == 0, "%s: %d nodes differ, first %d: %r vs %r" % (what, len(bad), bad[0], gpu_nodes[bad[0]], cpu_nodes[bad[0]])


REINSERT_DEFAULT = 16  # FLX_TUNE_BVH_REINSERT default


def build_both(tris, max_leaf=8, quality="fast", tri_cost=100, reinsert=None):
    from oracle.oracle_host import build_lbvh, build_ploc
    with CLContext(1024) as gpu:
        gpu.setTuning(bvh_tri_cost=tri_cost)
        if reinsert is not None:
            gpu.setTuning(bvh_reinsert=reinsert)
        nodes, idx, ms = gpu.buildBVH(tris, max_leaf, quality)
    if quality == "fast":
        cn, ci = build_lbvh(tris, max_leaf, tri_cost=tri_cost / 100.0)
    else:
        iterations = 0 if quality == "ploc" else (REINSERT_DEFAULT if reinsert is None else reinsert)
        cn, ci = build_ploc(tris, max_leaf, tri_cost=tri_cost / 100.0, reinsert=iterations)
    return nodes, idx, cn, ci, ms


QUALITIES = ["fast", "ploc", "ploc_opt"]


@pytest.mark.parametrize("quality", QUALITIES)
@pytest.mark.parametrize("max_leaf", [1, 8])
def test_builder_matches_the_cpu_restatement(max_leaf, quality):
    for name, scene in (("room", make_room_scene(materials="mixed", n_blobs=8)), ("teapot", teapot_scene())):
        nodes, idx, cn, ci, _ = build_both(scene.tris, max_leaf, quality)
        same_tree(nodes, idx, cn, ci, "%s max_leaf=%d %s" % (name, max_leaf, quality))
        validate_bvh(nodes, idx, scene.tris, max_leaf)


@pytest.mark.parametrize("quality", QUALITIES)
@pytest.mark.parametrize("tri_cost", [150, 200])
def test_builder_with_a_dearer_triangle_test_matches_the_cpu_restatement(quality, tri_cost):
    """FLX_TUNE_BVH_TRI_COST: the SAH collapse decision with a triangle test dearer than a box test (smaller leaves)"""
    scene = make_room_scene(materials="mixed", n_blobs=8)
    nodes, idx, cn, ci, _ = build_both(scene.tris, 8, quality, tri_cost)
    same_tree(nodes, idx, cn, ci, "room tri_cost=%d %s" % (tri_cost, quality))
    base = build_both(scene.tris, 8, quality)[0]
    assert len(nodes) >= len(base), "a dearer triangle test cannot make leaves larger"


@pytest.mark.parametrize("quality", QUALITIES)
def test_builder_edge_cases(quality):
    rng = np.random.default_rng(5)
    cases = {"one": rng.uniform(-1, 1, (1, 3, 3)), "two": rng.uniform(-1, 1, (2, 3, 3)), "identical": np.repeat(rng.uniform(-1, 1, (1, 3, 3)), 37, axis=0),
             "ragged": rng.uniform(-1, 1, (1001, 3, 3)) * 1e-3 + rng.uniform(-50, 50, (1001, 1, 3))}
    flat = rng.uniform(-1, 1, (200, 3, 3))
    flat[:, :, 2] = 0.25
    cases["flat"] = flat
    for name, pts in cases.items():
        tris = tri_soup(pts.astype(np.float32))
        nodes, idx, cn, ci, _ = build_both(tris, 4, quality)
        same_tree(nodes, idx, cn, ci, name + " " + quality)


@pytest.mark.parametrize("iterations", [1, 2, 5])
def test_reinsertion_iterations_match_the_cpu_restatement(iterations):
    """FLX_BVH_PLOC_OPT after 1, 2 and 5 iterations of the reinsertion pass (search, lock, guard, apply, refit): the same arrays as
    the sequential restatement, and a tree that is valid and no dearer than the plain PLOC one."""
    scene = SceneData.load_blob(scene_blob("conference"))
    nodes, idx, cn, ci, _ = build_both(scene.tris, 8, "ploc_opt", reinsert=iterations)
    same_tree(nodes, idx, cn, ci, "conference, %d reinsertion iterations" % iterations)
    depth, leaves, sah = validate_bvh(nodes, idx, scene.tris)
    base = build_both(scene.tris, 8, "ploc")
    assert sah < validate_bvh(base[0], base[1], scene.tris)[2], "reinsertion must lower the SAH cost on Conference"
    zero = build_both(scene.tris, 8, "ploc_opt", reinsert=0)
    same_tree(zero[0], zero[1], base[0], base[1], "0 iterations == FLX_BVH_PLOC")


def test_too_deep_optimised_tree_falls_back_to_the_plain_one():
    """The traversal stack holds 64 entries, so flx_build_bvh hands out no PLOC tree deeper than 62 levels.  Reinsertion can deepen a
    tree; past the limit the builder returns the PLOC tree it started from.  Reached here by lowering the limit (FLX_TUNE_BVH_DEPTH_LIMIT):
    Conference is 37 levels deep after PLOC and 38 after the reinsertion pass."""
    from fluctus_b200 import FluctusError
    from oracle.oracle_host import build_ploc
    scene = SceneData.load_blob(scene_blob("conference"))
    with CLContext(1024) as gpu:
        plain = gpu.buildBVH(scene.tris, 8, "ploc")
        assert validate_bvh(plain[0], plain[1], scene.tris)[0] == 37
        opt = gpu.buildBVH(scene.tris, 8, "ploc_opt")
        assert validate_bvh(opt[0], opt[1], scene.tris)[0] == 38
        gpu.setTuning(bvh_depth_limit=37)
        back = gpu.buildBVH(scene.tris, 8, "ploc_opt")
        same_tree(back[0], back[1], plain[0], plain[1], "PLOC_OPT past the depth limit == PLOC")
        cn, ci = build_ploc(scene.tris, 8, reinsert=REINSERT_DEFAULT, depth_limit=37)
        same_tree(back[0], back[1], cn, ci, "the restatement falls back the same way")
        gpu.setTuning(bvh_depth_limit=36)
        with pytest.raises(FluctusError, match="levels deep"):
            gpu.buildBVH(scene.tris, 8, "ploc_opt")
        with pytest.raises(FluctusError, match="levels deep"):
            gpu.buildBVH(scene.tris, 8, "ploc")


@pytest.mark.parametrize("quality", QUALITIES)
@pytest.mark.parametrize("name", ["conference", "country_kitchen"])
def test_builder_matches_on_reference_assets(name, quality):
    scene = SceneData.load_blob(scene_blob(name))
    nodes, idx, cn, ci, ms = build_both(scene.tris, 8, quality)
    same_tree(nodes, idx, cn, ci, name + " " + quality)
    assert ms < 50.0, "build took %.2f ms" % ms  # the reference's SBVH build takes seconds


@pytest.mark.parametrize("quality", QUALITIES)
def test_wavefront_through_the_built_tree_is_in_lockstep_with_the_oracle(quality):
    room = make_room_scene(materials="mixed", textured=True, n_blobs=8)
    W, H, N = 96, 64, 6144
    params = room_params(room, W, H, max_bounces=6, separate_queues=True)
    with CLContext(N) as gpu:
        nodes, idx, _ = gpu.buildBVH(room.tris, 8, quality)
        scene = SceneData(room.tris, idx, nodes, room.materials, room.tex_desc, room.tex_data)
        run_lockstep(gpu, oracle_ctx(N), scene, params, iterations=12)


def test_builder_rejects_bad_arguments():
    from fluctus_b200.clcontext import FluctusError
    tris = tri_soup(np.zeros((3, 3, 3), np.float32))
    with CLContext(64) as gpu:
        with pytest.raises(FluctusError):
            gpu.buildBVH(tris, max_leaf=0)
        with pytest.raises(FluctusError):
            gpu.buildBVH(tris, max_leaf=256)


@pytest.mark.parametrize("name", ["room", "teapot", "conference", "built"])
def test_device_repack_matches_the_host_repack(name):
    """flx_upload_scene makes the traversal layout (TNode / TTri) on the device (flx_bvh_repack.cuh); the host code it replaced
    stays as the checker behind a tuning knob: both must give the same bytes, on reference SBVHs and on a GPU-built tree."""
    if name == "room":
        scene = make_room_scene(materials="mixed", n_blobs=8)
    elif name == "teapot":
        scene = teapot_scene()
    else:
        scene = SceneData.load_blob(scene_blob("conference"))
    layouts = []
    for on_host in (1, 0):
        with CLContext(256) as gpu:
            gpu.setTuning(repack_on_host=on_host)
            if name == "built":
                nodes, idx, _ = gpu.buildBVH(scene.tris)
                sc = SceneData(scene.tris, idx, nodes, scene.materials, scene.tex_desc, scene.tex_data)
            else:
                sc = scene
            gpu.uploadSceneData(sc)
            layouts.append(gpu.readTraversalLayout())
    (hn, ht, hr), (dn, dt, dr) = layouts
    assert hr == dr and hn.shape == dn.shape and ht.shape == dt.shape
    assert np.array_equal(hn.view(np.uint32), dn.view(np.uint32)), "TNode records differ"
    assert np.array_equal(ht.view(np.uint32), dt.view(np.uint32)), "TTri records differ"


def test_device_repack_reports_bad_hierarchies():
    from fluctus_b200.clcontext import FluctusError
    room = make_room_scene(materials="diffuse")
    for what in ("leaf_range", "tri_index", "child_link"):
        nodes, indices = room.nodes.copy(), room.indices.copy()
        leaf = int(np.flatnonzero(nodes["nPrims"] > 0)[-1])
        inner = int(np.flatnonzero(nodes["nPrims"] == 0)[-1])
        if what == "leaf_range":
            nodes["link"][leaf] = len(indices)
        elif what == "tri_index":
            indices[nodes["link"][leaf]] = len(room.tris) + 7
        else:
            nodes["link"][inner] = len(nodes) + 3
        bad = SceneData(room.tris, indices, nodes, room.materials, room.tex_desc, room.tex_data)
        with CLContext(64) as gpu:
            with pytest.raises(FluctusError):
                gpu.uploadSceneData(bad)
