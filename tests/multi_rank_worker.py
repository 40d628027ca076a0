"""Worker run under ``python -m torch.distributed.run`` by test_multi_rank.py.

mode "host" (gloo, CPU): the host-side plumbing of ascent_b200.distributed against the oracle.
mode "gpu"  (nccl, one GPU per rank): sort-last path A (uint8 images, P2P fold, pulled and pushed) and
path B (float partials, P2P pull+merge+fold; ray layers) against the oracle at the same rank/block
layout ("parity is always vs the oracle at the same layout"), the opaque + volume scene,
and the failure behaviour of the collectives (a rank-local error releases the peers).
mode "gpu-shared" (gloo for the O(ranks) scalars, EVERY rank on cuda:0): the same checks with the
ranks' processes time-sharing one GPU -- the exchange arenas are mapped through CUDA IPC exactly as
across GPUs, so the cross-rank kernels run for real on a one-GPU box.
With VR_FULL_SIZE=1 (or 8 ranks) BASELINE configs 3 and 5 are also checked at their named sizes.
Exit code 0 = all checks passed."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from ascent_b200 import _lib, color_table, datasets, distributed as D  # noqa: E402


def scene(world, per_axis=2, n_block=12, az=25.0, W=320, H=200):
    from oracle import oracle as O
    doms = datasets.braid_uniform_blocks(n_block, per_axis, dtype=np.float32)
    bounds = [datasets.domain_bounds(d) for d in doms]
    gb = datasets.union_bounds(bounds)
    cam = O.camera_reset_to_bounds(gb)
    O.camera_azimuth(cam, az)
    O.camera_elevation(cam, az / 2.0)
    lut = color_table.parse_color_table({"name": "cool to warm", "control_points": [
        {"type": "alpha", "position": 0., "alpha": 0.},
        {"type": "alpha", "position": 1., "alpha": 1.}]}).corrected_opacity(100).lut()
    return dict(doms=doms, bounds=bounds, gb=gb, cam=cam, lut=lut, W=W, H=H,
                sample_dist=O.sample_distance(gb, 100))


def host_checks(rank, world):
    from oracle import oracle as O
    sc = scene(world)
    owners = D.assign_blocks(len(sc["doms"]), world)
    assert sorted(sum(owners, [])) == list(range(len(sc["doms"])))
    for mode in ("contiguous", "round_robin"):
        o = D.assign_blocks(13, world, mode)
        assert sorted(sum(o, [])) == list(range(13)) and max(map(len, o)) - min(map(len, o)) <= 1
    mine = owners[rank]
    # global scalar range / bounds == the serial values
    lmin = min(float(sc["doms"][i]["field"].min()) for i in mine)
    lmax = max(float(sc["doms"][i]["field"].max()) for i in mine)
    rmin, rmax = D.global_range(lmin, lmax, dist)
    assert rmin == min(float(d["field"].min()) for d in sc["doms"])
    assert rmax == max(float(d["field"].max()) for d in sc["doms"])
    gb = D.global_bounds([sc["bounds"][i] for i in mine], dist)
    assert np.array_equal(gb, sc["gb"])
    # path switch
    assert D.one_domain_per_rank(len(mine), dist) == (len(mine) == 1)
    assert D.one_domain_per_rank(1 if rank == 0 else 2, dist) is False
    # visibility order: identical integers to the oracle's serial DepthSort over (rank, domain)
    vis_all, vis_mine = D.global_visibility_order([sc["bounds"][i] for i in mine], sc["cam"], dist)
    ref, _ = O.visibility_order(np.array([sc["bounds"][i] for r in range(world) for i in owners[r]]), sc["cam"])
    assert np.array_equal(vis_all.reshape(-1), ref), (vis_all, ref)
    assert np.array_equal(vis_mine, ref.reshape(world, -1)[rank])
    # handle all-gather plumbing (bytes in, bytes out, rank order)
    hs = D.all_gather_array(np.full(64, rank, np.uint8), dist)
    assert hs.shape == (world, 64) and all((hs[r] == r).all() for r in range(world))
    # partial ownership formula of the merge kernel == oracle's RegularDecomposer restatement
    for (lo, hi) in [(0, 63999), (17, 1000), (5, 5), (100, 100 + world - 2)]:
        width = max((hi - lo + 1) // world, 1)
        for px in {lo, hi, (lo + hi) // 2, min(lo + width, hi), min(lo + width - 1, hi)}:
            own = [r for r in range(world)
                   if min(lo + width * r, hi + 1) <= px < (hi + 1 if r == world - 1 else min(lo + width * (r + 1), hi + 1))]
            assert own == [O.partial_owner(px, lo, hi, world)], (px, lo, hi, own)


def gpu_checks(rank, world):
    from oracle import oracle as O
    import scenes
    ctx = _lib.Context(torch.cuda.current_device())
    try:
        # ---------------- path B: 8 blocks over `world` ranks
        sc = scene(world)
        W, H = sc["W"], sc["H"]
        owners = D.assign_blocks(len(sc["doms"]), world)
        mine = owners[rank]
        rmin = min(float(d["field"].min()) for d in sc["doms"])
        rmax = max(float(d["field"].max()) for d in sc["doms"])
        ctx.set_tf(sc["lut"])
        for i in mine:
            ctx.block_from_domain(i, sc["doms"][i])
        D.connect(ctx, dist, W * H, W * H * len(mine))
        for rep in range(5):  # several frames: epoch parity / double buffering, moving camera
            cam = sc["cam"]
            if rep:
                O.camera_azimuth(cam, 7.0)
            layers = rep >= 3  # dense ray layers instead of lists
            fused = rep == 1 or layers  # pixels stored straight into rank 0's canvas, no list
            if layers:
                ctx.layers_begin(W, H)
                if rep == 4:   # the whole block loop in one call, launches overlapped
                    ctx.trace_blocks_to_layers(mine, cam, sc["sample_dist"], rmin, rmax, False)
                else:
                    for i in mine:
                        ctx.trace_to_layer(i, cam, sc["sample_dist"], rmin, rmax, False)
                ctx.comm_layers_composite_to_canvas(cam)
            else:
                ctx.partials_begin(W, H)
                for i in mine:
                    ctx.trace_to_partials(i, cam, sc["sample_dist"], rmin, rmax, False)
            if layers:
                pass
            elif fused:
                ctx.comm_composite_partials_to_canvas(cam)
            else:
                if rank == 0:
                    ctx.canvas_clear(W, H)
                ctx.comm_composite_partials()
            if rank == 0:
                got = ctx.partials_download() if not layers else np.zeros(0, O.PARTIAL_DTYPE)
                if not fused:
                    ctx.partials_to_canvas(cam)
                rgba, depth = ctx.canvas_download(W, H)
                o_rgba, o_depth = O.new_canvas(W, H)
                pl = [O.render_partials(scenes.oracle_block(sc["doms"][i]), cam, W, H, sc["lut"],
                                        sc["sample_dist"], rmin, rmax, o_depth)
                      for r in range(world) for i in owners[r]]
                ref = O.composite_partials(pl)
                a = np.sort(got, order="pixel_id")
                b = np.sort(ref, order="pixel_id")
                if fused:
                    assert a.size == 0
                else:
                    assert a.size == b.size and a.size > 1000, (a.size, b.size)
                    assert a.tobytes() == b.tobytes(), "path B: composited partials differ from the oracle"
                O.partials_to_canvas(ref, cam, W, H, o_rgba, o_depth)
                assert np.array_equal(rgba, o_rgba), "path B canvas differs (rep %d)" % rep
                cov = o_rgba[:, 3] > 0
                assert np.array_equal(depth[cov], o_depth[cov]) and cov.sum() > 1000
            elif not layers:
                assert ctx.partials_count() == 0  # root-only result
            dist.barrier()
        # ---------------- a rank-local error inside a collective must not hang the peers: the
        # last rank overflows max_partials -> it gets the error, everybody else is released, is told so by
        # its next synchronising call, and the following frame composites normally again
        if world > 1:
            cam = sc["cam"]
            ctx.partials_begin(W, H)
            reps = 1
            if rank == world - 1:
                area = sum(max(1, _lib.find_subset(cam, W, H, sc["bounds"][i])[2] *
                               _lib.find_subset(cam, W, H, sc["bounds"][i])[3]) for i in mine)
                reps = (W * H * max(len(o) for o in owners)) // area + 2  # connect() sized the arena for the largest rank
            for _ in range(reps):
                for i in mine:
                    ctx.trace_to_partials(i, cam, sc["sample_dist"], rmin, rmax, False)
            try:
                ctx.comm_composite_partials()
                failed_here = False
            except _lib.VRError as e:
                failed_here = True
                assert "max_partials" in str(e) and "aborted" in str(e), str(e)
            assert failed_here == (rank == world - 1)
            try:
                ctx.synchronize()
                told = False
            except _lib.VRError as e:
                told = True
                assert "aborted" in str(e) and "rank %d" % (world - 1) in str(e), str(e)
            assert told, "rank %d was not told that exchange was aborted" % rank
            dist.barrier()
            ctx.partials_begin(W, H)
            for i in mine:
                ctx.trace_to_partials(i, cam, sc["sample_dist"], rmin, rmax, False)
            ctx.comm_composite_partials_to_canvas(cam)
            if rank == 0:
                rgba, depth = ctx.canvas_download(W, H)
                o_rgba, o_depth = O.new_canvas(W, H)
                pl = [O.render_partials(scenes.oracle_block(sc["doms"][i]), cam, W, H, sc["lut"],
                                        sc["sample_dist"], rmin, rmax, o_depth)
                      for r in range(world) for i in owners[r]]
                O.partials_to_canvas(O.composite_partials(pl), cam, W, H, o_rgba, o_depth)
                assert np.array_equal(rgba, o_rgba), "path B canvas differs after an aborted exchange"
            else:
                ctx.synchronize()
            dist.barrier()
        for i in mine:
            ctx.block_free(i)

        # ---------------- path A: one block per rank
        doms = datasets.braid_uniform_blocks(14, 2, dtype=np.float32)[:world]
        bounds = [datasets.domain_bounds(d) for d in doms]
        gb = datasets.union_bounds(bounds)
        rmin = min(float(d["field"].min()) for d in doms)
        rmax = max(float(d["field"].max()) for d in doms)
        sd = O.sample_distance(gb, 100)
        ctx.block_from_domain(0, doms[rank])
        for rep, (az, el) in enumerate([(0., 0.), (160., 10.), (-70., 45.)]):
            cam = O.camera_reset_to_bounds(gb)
            O.camera_azimuth(cam, az)
            O.camera_elevation(cam, el)
            _, vis_mine = D.global_visibility_order([bounds[rank]], cam, dist)
            vis_all, _ = D.global_visibility_order([bounds[rank]], cam, dist)
            if rep == 0:   # the four separate calls ...
                ctx.canvas_clear(W, H)
                ctx.trace_to_canvas(0, cam, sd, rmin, rmax, False)
                ctx.image_from_canvas()
            else:          # ... or the fused frame kernel, outside of the block's rectangle unwritten
                ctx.trace_to_image(0, cam, W, H, sd, rmin, rmax, no_clear=(rep == 2))
            vo = np.ascontiguousarray(vis_all[:, 0], np.int32)
            if rep == 1:
                ctx.comm_composite_images_to_canvas(vo)   # ImageToCanvas folded into the exchange
            else:
                ctx.comm_composite_images(vo)
                if rank == 0:
                    ctx.image_result_to_canvas()
            if rank == 0:
                u8, d = ctx.image_result_download(W, H)
                can, cd = ctx.canvas_download(W, H)
                layers, depths = [], []
                for dom in doms:
                    r, dd = O.new_canvas(W, H)
                    O.render_to_canvas(scenes.oracle_block(dom), cam, W, H, sc["lut"], sd, rmin, rmax, r, dd)
                    q = O.image_init(r, dd, 0)
                    layers.append(q[0])
                    depths.append(q[1])
                order, _ = O.visibility_order(np.array(bounds), cam)
                ref, rd = O.ordered_composite(np.stack(layers), np.stack(depths), order)
                assert np.array_equal(u8, ref), "path A: composited uint8 image differs (rep %d)" % rep
                assert np.array_equal(d, rd, equal_nan=True), "path A: composited depth differs (rep %d)" % rep
                o_can, o_cd = O.image_to_canvas(ref, rd)
                assert np.array_equal(can, o_can) and np.array_equal(cd, o_cd, equal_nan=True), "path A canvas (rep %d)" % rep
                cov = ref[:, 3] > 0
                assert cov.sum() > 1000
            else:
                ctx.synchronize()
            dist.barrier()

        # ---------------- path A, renders of a batch software-pipelined (VR_FRAME_AHEAD): trace(k+1) is
        # issued before exchange(k); every exchanged frame must still be the oracle's frame k
        def cam_of(k):
            c = O.camera_reset_to_bounds(gb)
            O.camera_azimuth(c, 15.0 + 40.0 * k)
            O.camera_elevation(c, 5.0 * k)
            return c

        def oracle_frame(c):
            layers, depths = [], []
            for dom in doms:
                r, dd = O.new_canvas(W, H)
                O.render_to_canvas(scenes.oracle_block(dom), c, W, H, sc["lut"], sd, rmin, rmax, r, dd)
                q = O.image_init(r, dd, 0)
                layers.append(q[0])
                depths.append(q[1])
            order, _ = O.visibility_order(np.array(bounds), c)
            return O.ordered_composite(np.stack(layers), np.stack(depths), order)

        # frames 3.. are PUSHED (VR_FRAME_PUSH): the sampler stores every pixel into its owner's receive
        # slot, the fold reads local memory; frame 5 mixes pushed (even ranks) and pulled (odd ranks) images
        def pushed(k):
            return k in (3, 4, 6, 7, 8) or (k == 5 and rank % 2 == 0)

        n_frames = 9
        ctx.trace_to_image(0, cam_of(0), W, H, sd, rmin, rmax, no_clear=True)
        for k in range(n_frames):
            if k + 1 < n_frames:
                ctx.trace_to_image(0, cam_of(k + 1), W, H, sd, rmin, rmax, no_clear=True, ahead=True,
                                   push=pushed(k + 1))
                if k == 0:
                    try:   # only one frame may be ahead
                        ctx.trace_to_image(0, cam_of(k + 2), W, H, sd, rmin, rmax, no_clear=True, ahead=True)
                        raise AssertionError("a second frame ahead must be refused")
                    except _lib.VRError:
                        pass
            vis_all, _ = D.global_visibility_order([bounds[rank]], cam_of(k), dist)
            ctx.comm_composite_images_to_canvas(np.ascontiguousarray(vis_all[:, 0], np.int32))
            if rank == 0:
                u8, d = ctx.image_result_download(W, H)
                ref, rd = oracle_frame(cam_of(k))
                assert np.array_equal(u8, ref), "pipelined path A: frame %d differs" % k
                assert np.array_equal(d, rd, equal_nan=True), "pipelined path A: depth of frame %d differs" % k
                can, cd = ctx.canvas_download(W, H)
                o_can, o_cd = O.image_to_canvas(ref, rd)
                assert np.array_equal(can, o_can) and np.array_equal(cd, o_cd, equal_nan=True)
            else:
                ctx.synchronize()
            dist.barrier()

        # ---------------- opaque surfaces + volume (Scene::Render passes 1 and 2, Scene.cpp:160-214):
        # every rank holds an opaque canvas (here: synthetic fragments, incl. cross-rank depth ties and
        # far fragments) -> z-buffer composite to rank 0 -> ImageToCanvas -> SynchDepths -> volume pass
        # clamped at the synchronised depth, blended over each rank's canvas -> vis-order composite
        cam = O.camera_reset_to_bounds(gb)
        O.camera_azimuth(cam, 20.0)
        vis_all, _ = D.global_visibility_order([bounds[rank]], cam, dist)
        vo = np.ascontiguousarray(vis_all[:, 0], np.int32)
        for rep in range(3):
            def opaque(r):
                g = np.random.default_rng(100 * rep + r)
                col = np.zeros((H, W, 4), np.float32)
                dep = np.full((H, W), 1.001, np.float32)
                for _ in range(6):
                    x0, y0 = int(g.integers(0, W - 40)), int(g.integers(0, H - 30))
                    w, h = int(g.integers(20, 160)), int(g.integers(20, 120))
                    col[y0:y0 + h, x0:x0 + w] = [g.random(), g.random(), g.random(), 1.0]
                    # a small set of depth values so that fragments of different ranks tie exactly
                    dep[y0:y0 + h, x0:x0 + w] = np.float32(g.choice([0.80, 0.90, 0.95, 0.97, 0.99]))
                dep[:8, :] = np.float32(1.5)  # beyond the far plane: never replaces
                col[:8, :] = [0.3, 0.3, 0.3, 1.0]
                return col.reshape(-1, 4), dep.reshape(-1)
            my_col, my_dep = opaque(rank)
            ctx.canvas_upload(W, H, my_col, my_dep)
            ctx.image_from_canvas()
            ctx.comm_composite_zbuffer()
            if rank == 0:
                ctx.image_result_to_canvas()
                zu8, zd = ctx.image_result_download(W, H)
            ctx.comm_sync_depths()
            ctx.trace_to_canvas(0, cam, sd, rmin, rmax, True)
            sync_can, sync_depth = ctx.canvas_download(W, H)
            ctx.image_from_canvas()
            ctx.comm_composite_images(vo)
            # ---- oracle, every rank computes the whole thing (small) and checks its own part
            cols, deps = zip(*[opaque(r) for r in range(world)])
            # RadixKCompositor: equal-depth fragments of different ranks resolve in the tree's visiting order
            qs = [O.image_init(cols[r], deps[r], 0) for r in range(world)]
            front, fd = O.radixk_zbuffer(np.stack([q[0] for q in qs]), np.stack([q[1] for q in qs]), W, H)
            root_can, root_depth = O.image_to_canvas(front, fd)
            layers, depths = [], []
            for r in range(world):
                c_r = (root_can if r == 0 else cols[r]).copy()
                d_r = root_depth.copy()   # SynchDepths: rank 0's depth everywhere
                O.render_to_canvas(scenes.oracle_block(doms[r]), cam, W, H, sc["lut"], sd, rmin, rmax, c_r, d_r,
                                   use_depth=True)
                if r == rank:
                    assert np.array_equal(sync_can, c_r), "opaque+volume: canvas after the volume pass (rep %d)" % rep
                    assert np.array_equal(sync_depth, d_r, equal_nan=True), "opaque+volume: depth (rep %d)" % rep
                q = O.image_init(c_r, d_r, 0)
                layers.append(q[0])
                depths.append(q[1])
            if rank == 0:
                assert np.array_equal(zu8, front) and np.array_equal(zd, fd), "z-buffer composite differs (rep %d)" % rep
                assert (front[:, 3] == 255).sum() > 1000
                order, _ = O.visibility_order(np.array(bounds), cam)
                ref, rd = O.ordered_composite(np.stack(layers), np.stack(depths), order)
                u8, d = ctx.image_result_download(W, H)
                assert np.array_equal(u8, ref) and np.array_equal(d, rd, equal_nan=True), \
                    "opaque+volume: final composite differs (rep %d)" % rep
            else:
                ctx.synchronize()
            dist.barrier()
    finally:
        ctx.close()


def main():
    mode = sys.argv[1]
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    if mode == "host":
        dist.init_process_group("gloo")
        host_checks(rank, world)
    elif mode == "gpu-shared":
        torch.cuda.set_device(0)
        dist.init_process_group("gloo")
        gpu_checks(rank, world)
    else:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
        dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
        gpu_checks(rank, world)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("multi-rank %s checks OK (world=%d)" % (mode, world))


if __name__ == "__main__":
    main()
