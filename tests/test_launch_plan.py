"""Host logic of the launch plan (no GPU): sample slices per pixel and the longest-first launch order of a device's
8x4-pixel tiles (ptcuda.cu: plan_slices, plan_tile_order), through the debug hook ptc_debug_launch_plan."""
import numpy as np
import pytest

from pathtracer_ocl_b200 import scene as S, trace as T


def tile_centres(order, width, rows):
    tiles_x = (width + 7) // 8
    ty, tx = np.divmod(order, tiles_x)
    return tx * 8 + 4, np.asarray(rows)[np.minimum(ty * 4 + 1, len(rows) - 1)]


def test_teapot_launches_the_tiles_that_look_at_the_mesh_first():
    W, H = 640, 480
    sc = S.build_scene("teapot", W, H)
    slices, order = T.debug_launch_plan(sc, 2048)
    n_tiles = ((W + 7) // 8) * ((H + 3) // 4)
    assert slices == 8 and len(order) == n_tiles
    assert np.array_equal(np.sort(order), np.arange(n_tiles))           # a permutation: every tile rendered exactly once
    # heavy tiles form a prefix in frame order, light tiles the rest in frame order: one descent in the tile index
    descents = np.flatnonzero(np.diff(order) < 0)
    assert len(descents) == 1
    n_heavy = descents[0] + 1
    assert 0.02 * n_tiles < n_heavy < 0.5 * n_tiles
    x, y = tile_centres(order, W, np.arange(H))
    hx, hy = x[:n_heavy], y[:n_heavy]
    lx, ly = x[n_heavy:], y[n_heavy:]
    # the heavy set is the bounding rectangle of the mesh on screen (plus a margin): no light tile lies inside it
    assert not ((lx > hx.min()) & (lx < hx.max()) & (ly > hy.min()) & (ly < hy.max())).any()
    # the teapot stands on the floor in the middle of the room (translate(0, -0.4, 0), scenes/teapot.go): its tiles are
    # horizontally centred, below the image centre, and the image corners are light
    assert abs(hx.mean() - W / 2) < W / 10 and H / 2 < hy.mean() < 0.9 * H
    corners = {0, (W + 7) // 8 - 1, n_tiles - 1}
    assert corners <= set(order[n_heavy:].tolist())


@pytest.mark.parametrize("name", ["gopher", "cubemap", "transparent_teapot"])
def test_every_mesh_scene_gets_a_permutation(name):
    """The geometric estimate may decline (a mesh whose screen rectangle covers the frame, a corner behind the camera):
    then the order is the identity until the cost probe has measured the tiles."""
    W, H = 320, 240
    sc = S.build_scene(name, W, H, tex_scale=16)
    slices, order = T.debug_launch_plan(sc, 2048)
    n_tiles = ((W + 7) // 8) * ((H + 3) // 4)
    assert np.array_equal(np.sort(order), np.arange(n_tiles))


def test_scenes_without_meshes_keep_frame_order_and_small_frames_get_more_slices():
    sc = S.build_scene("reference", 1280, 960, 0.15, 1.6)
    slices, order = T.debug_launch_plan(sc, 2048)
    assert slices == 8 and np.array_equal(order, np.arange(len(order)))
    assert T.debug_launch_plan(S.build_scene("reference", 1280, 960, 0.15, 1.6), 2048, shard_index=3, shard_count=8)[0] == 8
    assert T.debug_launch_plan(S.build_scene("reference", 96, 72), 2048)[0] == 32          # 6912 pixels cannot fill 148 SMs
    assert T.debug_launch_plan(S.build_scene("reference", 96, 72), 5)[0] == 4              # never more slices than samples
    assert T.debug_launch_plan(S.build_scene("default", 16, 12), 1)[0] == 1


def test_shards_partition_the_tiles_and_order_their_own_rows():
    W, H, world = 320, 240, 4
    sc = S.build_scene("teapot", W, H)
    seen_rows = []
    for r in range(world):
        rows = T.plan_rows(H, r, world)
        slices, order = T.debug_launch_plan(sc, 256, shard_index=r, shard_count=world)
        n_tiles = ((W + 7) // 8) * ((len(rows) + 3) // 4)
        assert len(order) == n_tiles and np.array_equal(np.sort(order), np.arange(n_tiles))
        x, y = tile_centres(order, W, rows)
        first = order[: max(1, n_tiles // 20)]
        fx, fy = tile_centres(first, W, rows)
        assert abs(fx.mean() - W / 2) < W / 4 and abs(fy.mean() - H / 2) < H / 3        # the teapot's tiles lead
        seen_rows.extend(rows.tolist())
    assert sorted(seen_rows) == list(range(H))
